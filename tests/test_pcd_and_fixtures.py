"""PCD I/O and the weak fixtures the reference ships: relations between its .pcd files."""
import hashlib
import json
import os

import numpy as np

from conftest import ROOT, cloud_path
from realtime_robot_b200.pcd import read_pcd_xyz, to_xyz1, write_pcd_xyz


def test_manifest_matches_files():
    man = json.load(open(os.path.join(ROOT, "data", "clouds", "MANIFEST.json")))
    assert man["chair1"]["points"] == 1818 and man["mcloud"]["points"] == 1909 and man["sofa"]["points"] == 23172
    for name, rec in man.items():
        xyz = read_pcd_xyz(cloud_path(name))
        assert len(xyz) == rec["points"]
        assert hashlib.sha1(xyz.tobytes()).hexdigest() == rec["xyz_sha1"]


def test_roundtrip_binary_and_ascii(tmp_path):
    rng = np.random.default_rng(0)
    xyz = (rng.standard_normal((257, 3)) * 3).astype(np.float32)
    for binary in (True, False):
        p = tmp_path / f"c{int(binary)}.pcd"
        write_pcd_xyz(p, xyz, binary=binary)
        back = read_pcd_xyz(p)
        assert back.dtype == np.float32 and np.array_equal(back, xyz)


def test_reads_extra_fields(tmp_path):
    # binary xyz + rgb (chair1-style) and ascii xyz + normals + curvature (Chair_025-style)
    n = 5
    rec = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgb", "<f4")])
    rec["x"], rec["y"], rec["z"] = np.arange(n), np.arange(n) * 2, np.arange(n) * 3
    p = tmp_path / "rgb.pcd"
    with open(p, "wb") as f:
        f.write(b"# .PCD v0.7\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 5\nHEIGHT 1\n"
                b"VIEWPOINT 0 0 0 1 0 0 0\nPOINTS 5\nDATA binary\n")
        f.write(rec.tobytes())
    got = read_pcd_xyz(p)
    assert np.array_equal(got[:, 1], np.arange(n) * 2)
    p2 = tmp_path / "nrm.pcd"
    with open(p2, "w") as f:
        f.write("VERSION 0.7\nFIELDS x y z normal_x normal_y normal_z curvature\nSIZE 4 4 4 4 4 4 4\nTYPE F F F F F F F\n"
                "COUNT 1 1 1 1 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA ascii\n1 2 3 0 0 1 0.5\n4 5 6 0 1 0 -4.3e8\n")
    assert np.array_equal(read_pcd_xyz(p2), np.array([[1, 2, 3], [4, 5, 6]], np.float32))


def test_mcloud_is_5_translated():
    # mcloud.pcd == 5.pcd + (0.8, 0.8, 0): a known pure-translation pair shipped by the reference
    a, b = read_pcd_xyz(cloud_path("5")), read_pcd_xyz(cloud_path("mcloud"))
    assert np.abs(b - (a + np.array([0.8, 0.8, 0.0], np.float32))).max() < 2e-7


def test_70761_c_is_translated():
    # the commented generator at function.h:131-143: x + 1, y + 0.5
    a, b = read_pcd_xyz(cloud_path("70761")), read_pcd_xyz(cloud_path("70761_c"))
    assert np.abs(b - (a + np.array([1.0, 0.5, 0.0], np.float32))).max() < 2e-6


def test_xyz1_layout():
    p = to_xyz1(np.array([[1, 2, 3]], np.float32))
    assert p.shape == (1, 4) and p[0, 3] == 1.0 and p.nbytes == 16
