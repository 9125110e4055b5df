"""PCD v0.7 reader / writer (ascii + binary; binary_compressed through librtr.so), xyz extraction.

Harness-side mirror of the two PCL I/O calls the reference path makes:
``pcl::io::loadPCDFile`` (RealTimeRobot/RealTimeRobot.cpp:34-35, scan_point.h:62) and
``pcl::io::savePCDFileASCII`` (RealTimeRobot.cpp:108-109, function.h:126-127).
Only the x/y/z fields are kept, exactly as the reference loads every file into a
``pcl::PointCloud<pcl::PointXYZ>``.  The C++ host library has its own reader
(realtime_robot_b200/host/pcd_io.h); this one feeds the tests and the bench.
"""
from __future__ import annotations

import io
import numpy as np

_NP = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 1): "u1", ("U", 2): "<u2", ("U", 4): "<u4",
       ("U", 8): "<u8", ("I", 1): "i1", ("I", 2): "<i2", ("I", 4): "<i4", ("I", 8): "<i8"}


def _parse_header(f):
    hdr = {}
    while True:
        line = f.readline()
        if not line:
            raise ValueError("PCD: unexpected end of header")
        s = line.decode("ascii", "replace").strip()
        if not s or s.startswith("#"):
            continue
        key, _, rest = s.partition(" ")
        hdr[key.upper()] = rest.split()
        if key.upper() == "DATA":
            break
    return hdr


def read_pcd_xyz(path) -> np.ndarray:
    """Return an (N, 3) float32 array of the x, y, z fields of a PCD v0.7 file."""
    with open(path, "rb") as f:
        hdr = _parse_header(f)
        fields = hdr["FIELDS"]
        sizes = [int(v) for v in hdr["SIZE"]]
        types = hdr["TYPE"]
        counts = [int(v) for v in hdr.get("COUNT", ["1"] * len(fields))]
        npts = int(hdr["POINTS"][0]) if "POINTS" in hdr else int(hdr["WIDTH"][0]) * int(hdr["HEIGHT"][0])
        mode = hdr["DATA"][0].lower()
        for name in ("x", "y", "z"):
            if name not in fields:
                raise ValueError(f"PCD: field {name} missing in {path}")
        if mode == "ascii":
            cols, c = {}, 0
            for name, cnt in zip(fields, counts):
                cols[name] = c
                c += cnt
            data = np.loadtxt(io.BytesIO(f.read()), dtype=np.float64, ndmin=2)
            if data.shape[0] != npts:
                raise ValueError(f"PCD: {path}: expected {npts} rows, got {data.shape[0]}")
            return np.ascontiguousarray(data[:, [cols["x"], cols["y"], cols["z"]]].astype(np.float32))
        if mode == "binary":
            dt = np.dtype([(n, _NP[(t, s)], (c,)) if c != 1 else (n, _NP[(t, s)])
                           for n, s, t, c in zip(fields, sizes, types, counts)])
            rec = np.frombuffer(f.read(npts * dt.itemsize), dtype=dt, count=npts)
            return np.ascontiguousarray(
                np.stack([rec["x"], rec["y"], rec["z"]], axis=1).astype(np.float32))
        if mode == "binary_compressed":          # LZF codec lives in the library (csrc/pcd_io.cu); host-only, no GPU needed
            from . import api
            return np.ascontiguousarray(api.read_pcd(str(path))[:, :3])
        raise ValueError(f"PCD: DATA {mode} not supported")


def _header(n: int, mode: str) -> str:
    return ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\n"
            "TYPE F F F\nCOUNT 1 1 1\n"
            f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {mode}\n")


def write_pcd_xyz(path, xyz: np.ndarray, binary: bool = True) -> None:
    """Write an (N, 3) float32 array as an unorganised xyz PCD v0.7 file."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    with open(path, "wb") as f:
        f.write(_header(len(xyz), "binary" if binary else "ascii").encode("ascii"))
        if binary:
            f.write(xyz.astype("<f4").tobytes())
        else:
            # %.9g round-trips every float32 (savePCDFileASCII prints 8 significant digits)
            for p in xyz:
                f.write(("%.9g %.9g %.9g\n" % (p[0], p[1], p[2])).encode("ascii"))


def to_xyz1(xyz: np.ndarray) -> np.ndarray:
    """(N,3) -> (N,4) float32 with pad = 1.0f: the 16-byte pcl::PointXYZ layout."""
    xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
    out = np.ones((len(xyz), 4), dtype=np.float32)
    out[:, :3] = xyz
    return out
