"""CPU restatement of the PCD v0.7 container for the I/O row of the scope table: header, DATA ascii /
binary / binary_compressed, and the LZF stream PCL uses for the latter.  TEST INFRASTRUCTURE: only tests/ may import this;
the product reader / writer is realtime_robot_b200/csrc/pcd_io.cu.  **Parity unpinned**: the reference ships no
binary_compressed file and PCL is not installed, so the compressed layout (two little-endian uint32 sizes, then an LZF
stream of the field-major arrays) and the liblzf control-byte format are restated from their published descriptions; the
reference's own ascii and binary .pcd files pin the other two modes (tests/test_pcd_io.py reads them with both readers)."""
import numpy as np

_NP = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 1): "u1", ("U", 2): "<u2", ("U", 4): "<u4", ("I", 1): "i1", ("I", 2): "<i2", ("I", 4): "<i4"}


def lzf_decompress(src: bytes, out_len: int) -> bytes:
    """liblzf stream: ctrl < 32 -> ctrl + 1 literals; else length (ctrl >> 5) + 2 (+ next byte when the 3 bits are 7),
    distance ((ctrl & 31) << 8 | next byte) + 1, byte-wise copy (may overlap)."""
    out = bytearray()
    ip, n = 0, len(src)
    while ip < n:
        ctrl = src[ip]; ip += 1
        if ctrl < 32:
            out += src[ip:ip + ctrl + 1]; ip += ctrl + 1
        else:
            ln = ctrl >> 5
            if ln == 7:
                ln += src[ip]; ip += 1
            dist = (((ctrl & 31) << 8) | src[ip]) + 1; ip += 1
            ln += 2
            start = len(out) - dist
            if start < 0:
                raise ValueError("LZF: reference before the start of the output")
            if dist >= ln:
                out += out[start:start + ln]
            else:
                for k in range(ln):
                    out.append(out[start + k])
    if len(out) != out_len:
        raise ValueError(f"LZF: decoded {len(out)} bytes, header says {out_len}")
    return bytes(out)


def lzf_compress(src: bytes) -> bytes:
    """A deliberately different encoder from the product's (dictionary of the last position of every 3-byte string,
    greedy, plus forced long runs) so that decoder tests do not depend on one encoder's habits."""
    out = bytearray()
    lits = bytearray()
    last = {}
    ip, n = 0, len(src)

    def flush():
        nonlocal lits
        for k in range(0, len(lits), 32):
            chunk = lits[k:k + 32]
            out.append(len(chunk) - 1)
            out.extend(chunk)
        lits = bytearray()

    while ip < n:
        key = src[ip:ip + 3]
        r = last.get(key) if len(key) == 3 else None
        if len(key) == 3:
            last[key] = ip
        if r is not None and ip - r <= 8192:
            ln = 3
            while ln < 264 and ip + ln < n and src[r + ln] == src[ip + ln]:
                ln += 1
            flush()
            off, l = ip - r - 1, ln - 2
            if l < 7:
                out.append((off >> 8) + (l << 5))
            else:
                out.append((off >> 8) + (7 << 5)); out.append(l - 7)
            out.append(off & 0xff)
            ip += ln
        else:
            lits.append(src[ip]); ip += 1
    flush()
    return bytes(out)


def _header(f):
    hdr = {}
    while True:
        line = f.readline()
        if not line:
            raise ValueError("PCD: no DATA line")
        s = line.decode("ascii", "replace").strip()
        if not s or s.startswith("#"):
            continue
        k, _, rest = s.partition(" ")
        hdr[k.upper()] = rest.split()
        if k.upper() == "DATA":
            return hdr


def read_xyz(path) -> np.ndarray:
    """(N, 3) float32 from a PCD v0.7 file in any of the three DATA modes."""
    with open(path, "rb") as f:
        hdr = _header(f)
        fields, sizes, types = hdr["FIELDS"], [int(v) for v in hdr["SIZE"]], hdr["TYPE"]
        counts = [int(v) for v in hdr.get("COUNT", ["1"] * len(fields))]
        n = int(hdr["POINTS"][0]) if "POINTS" in hdr else int(hdr["WIDTH"][0]) * int(hdr["HEIGHT"][0])
        mode = hdr["DATA"][0]
        body = f.read()
    if mode == "ascii":
        col, c = {}, 0
        for name, cnt in zip(fields, counts):
            col[name] = c; c += cnt
        rows = [ln.split() for ln in body.decode("ascii").splitlines() if ln.strip()][:n]
        return np.array([[np.float32(r[col[a]]) for a in "xyz"] for r in rows], dtype=np.float32).reshape(n, 3)
    if mode == "binary":
        dt = np.dtype([(nm, _NP[(t, s)], (c,)) if c != 1 else (nm, _NP[(t, s)]) for nm, s, t, c in zip(fields, sizes, types, counts)])
        rec = np.frombuffer(body[:n * dt.itemsize], dtype=dt, count=n)
        return np.stack([rec["x"], rec["y"], rec["z"]], 1).astype(np.float32)
    if mode == "binary_compressed":
        csize, usize = np.frombuffer(body[:8], dtype="<u4")
        raw = lzf_decompress(body[8:8 + int(csize)], int(usize))
        out, off = {}, 0
        for nm, s, t, c in zip(fields, sizes, types, counts):
            if nm in ("x", "y", "z"):
                out[nm] = np.frombuffer(raw[off:off + n * s * c], dtype=_NP[(t, s)], count=n)
            off += n * s * c
        return np.stack([out["x"], out["y"], out["z"]], 1).astype(np.float32)
    raise ValueError(mode)


def write_xyz(path, xyz, mode: str, extra_field: bool = False) -> None:
    """Writer for test fixtures; `extra_field` interleaves an rgb field (as the reference's binary files have) so readers
    must honour FIELDS / SIZE instead of assuming xyz-only records."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    n = len(xyz)
    fields = "x y z rgb" if extra_field else "x y z"
    k = 4 if extra_field else 3
    hdr = (f"# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS {fields}\nSIZE {' '.join(['4'] * k)}\n"
           f"TYPE {' '.join(['F'] * k)}\nCOUNT {' '.join(['1'] * k)}\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {mode}\n")
    cols = [xyz[:, 0], xyz[:, 1], xyz[:, 2]] + ([np.full(n, 4.2e-39, np.float32)] if extra_field else [])
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        if mode == "ascii":
            for i in range(n):
                f.write((" ".join("%.9g" % float(c[i]) for c in cols) + "\n").encode("ascii"))
        elif mode == "binary":
            f.write(np.stack(cols, 1).astype("<f4").tobytes())
        else:
            raw = b"".join(np.ascontiguousarray(c, dtype="<f4").tobytes() for c in cols)
            comp = lzf_compress(raw)
            f.write(np.array([len(comp), len(raw)], dtype="<u4").tobytes())
            f.write(comp)
