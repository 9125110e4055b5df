#!/usr/bin/env python3
"""Derive the benchmark / test clouds from the reference's .pcd fixtures.

Run in the build container only (it reads /root/reference, which does not exist on the
GPU box).  Every reference cloud is loaded the way the reference loads it — x, y, z only,
into a PointXYZ cloud (RealTimeRobot.cpp:32-35) — and re-written with THIS repo's writer as
an unorganised binary xyz PCD under data/clouds/.  The rgb / rgba / normal / curvature fields
are dropped, so no file is a byte copy of a reference file.  Byte-identical duplicates in the
reference (desk2 == desk1, 8 == chair2, "11351 (1)" == 11351) are imported once.
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from realtime_robot_b200.pcd import read_pcd_xyz, write_pcd_xyz  # noqa: E402

REF = "/root/reference/RealTimeRobot"
OUT = os.path.join(os.path.dirname(__file__), "..", "data", "clouds")
NAMES = ["chair1", "chair2", "chair4", "desk1", "desk3", "sofa", "ground", "Chair_025", "mcloud",
         "5", "6", "T0_m8081", "T0_m8111", "T0_m8241", "T0_m8261", "11351", "39851", "46631",
         "70081", "70761", "70761_c"]


def main():
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for name in NAMES:
        xyz = read_pcd_xyz(os.path.join(REF, name + ".pcd"))
        dst = os.path.join(OUT, name + ".pcd")
        write_pcd_xyz(dst, xyz, binary=True)
        manifest[name] = {"points": int(len(xyz)),
                          "xyz_sha1": hashlib.sha1(xyz.tobytes()).hexdigest(),
                          "min": [float(v) for v in xyz.min(0)], "max": [float(v) for v in xyz.max(0)]}
        print(f"{name:12s} {len(xyz):6d} pts")
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
