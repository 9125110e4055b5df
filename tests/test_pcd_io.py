"""PCD v0.7 I/O of librtr.so (csrc/pcd_io.cu: ascii, binary, binary_compressed; host side, no GPU needed) against the CPU
restatement oracle/pcd_ref.py and against the reference's own .pcd files (ascii xyz, binary xyz+rgb)."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import pcd_ref
from realtime_robot_b200 import api
from realtime_robot_b200.pcd import read_pcd_xyz


def _cloud(n, seed=0):
    rng = np.random.default_rng(seed)
    p = np.ones((n, 4), np.float32)
    p[:, :3] = (rng.standard_normal((n, 3)) * [3.0, 1.0, 0.2] + [0.5, -2.0, 10.0]).astype(np.float32)
    return p


def test_lzf_restatement_round_trip():
    rng = np.random.default_rng(1)
    for raw in (b"", b"a", b"abcabcabcabcabcabc" * 40, bytes(1000), rng.integers(0, 256, 5000, dtype=np.uint8).tobytes(),
                np.repeat(rng.integers(0, 4, 700, dtype=np.uint8), 9).tobytes()):
        assert pcd_ref.lzf_decompress(pcd_ref.lzf_compress(raw), len(raw)) == raw
    # hand-made stream: 3 literals "abc", then a back reference of length 9 at distance 3 (overlapping copy)
    assert pcd_ref.lzf_decompress(bytes([2]) + b"abc" + bytes([(7 << 5) | 0, 0, 2]), 12) == b"abc" * 4


@pytest.mark.parametrize("mode", [api.PCD_ASCII, api.PCD_BINARY, api.PCD_BINARY_COMPRESSED])
@pytest.mark.parametrize("n", [0, 1, 7, 1000, 70000])
def test_library_writer_read_by_the_restatement(tmp_path, mode, n):
    p = _cloud(n, n)
    if n >= 7:
        p[3, :3] = [0.0, -0.0, 1e-38]            # zero, negative zero, denormal-range value
        p[5, :3] = [1e30, -1e-30, 123456.789]
    path = str(tmp_path / "w.pcd")
    api.write_pcd(path, p, mode)
    assert api.pcd_info(path) == (n, mode)
    back = pcd_ref.read_xyz(path)
    if mode == api.PCD_ASCII:                     # 8 significant digits, like savePCDFileASCII: not bit-exact by design
        assert np.allclose(back, p[:, :3], rtol=1.3e-7, atol=1e-37)       # within one float ulp
    else:
        assert np.array_equal(back.view(np.uint32), p[:, :3].view(np.uint32))
    again = api.read_pcd(path)                    # and the library reads its own files
    assert np.array_equal(again[:, :3].view(np.uint32), back.view(np.uint32)) and np.all(again[:, 3] == 1.0)


@pytest.mark.parametrize("mode", ["ascii", "binary", "binary_compressed"])
@pytest.mark.parametrize("extra", [False, True])
def test_library_reader_on_restatement_files(tmp_path, mode, extra):
    p = _cloud(4321, 3)
    path = str(tmp_path / "r.pcd")
    pcd_ref.write_xyz(path, p[:, :3], mode, extra_field=extra)
    got = api.read_pcd(path)
    assert np.array_equal(got[:, :3].view(np.uint32), p[:, :3].view(np.uint32))      # %.9g round-trips float32
    assert np.array_equal(read_pcd_xyz(path), p[:, :3]) if mode != "binary_compressed" else True


def test_reader_matches_on_every_repo_cloud():
    files = sorted(glob.glob(os.path.join(ROOT, "data", "clouds", "*.pcd")))
    assert len(files) >= 10
    for f in files:
        a, b = api.read_pcd(f), pcd_ref.read_xyz(f)
        assert a.shape[0] == b.shape[0] and np.array_equal(a[:, :3].view(np.uint32), b.view(np.uint32)), f


def test_ascii_writer_prints_like_savePCDFileASCII(tmp_path):
    p = np.array([[0.1, -2.5, 3.0, 1], [1e-5, 123456792.0, -0.000123456789, 1], [np.nan, np.inf, -np.inf, 1]], np.float32)
    path = str(tmp_path / "a.pcd")
    api.write_pcd(path, p, api.PCD_ASCII)
    lines = open(path).read().splitlines()
    assert lines[-3:] == ["0.1 -2.5 3", "9.9999997e-06 1.2345679e+08 -0.00012345679", "nan inf -inf"]   # "%.8g"
    assert lines[:2] == ["# .PCD v0.7 - Point Cloud Data file format", "VERSION 0.7"] and lines[-4] == "DATA ascii"


def test_errors(tmp_path):
    with pytest.raises(Exception):
        api.read_pcd(str(tmp_path / "missing.pcd"))
    bad = tmp_path / "bad.pcd"
    bad.write_text("VERSION 0.7\nFIELDS x y\nSIZE 4 4\nTYPE F F\nCOUNT 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA ascii\n1 2\n")
    with pytest.raises(Exception):
        api.read_pcd(str(bad))
    trunc = tmp_path / "trunc.pcd"
    trunc.write_bytes(b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 4\nHEIGHT 1\nPOINTS 4\nDATA binary\n" + bytes(20))
    with pytest.raises(Exception):
        api.read_pcd(str(trunc))
    corrupt = tmp_path / "c.pcd"
    corrupt.write_bytes(b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA binary_compressed\n"
                        + np.array([3, 24], "<u4").tobytes() + bytes([0xe0, 0x10, 0x05]))
    with pytest.raises(Exception):
        api.read_pcd(str(corrupt))
