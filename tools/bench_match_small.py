#!/usr/bin/env python3
"""Descriptor matching of the bench batch on its own: the real FPFH rows of the 8 repo models (70 k rows) against mcloud's
(1909 rows), k = 5, through rtr_match_features_raw.  Prints the device time of the search and the rows the certificate sent to
the exact kernel — run with RTR_MATCH_KEEP=8 / 16 and RTR_MATCH_TC=0 to compare the variants."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from realtime_robot_b200 import api  # noqa: E402
from realtime_robot_b200.pcd import read_pcd_xyz, to_xyz1  # noqa: E402

MODELS = ["chair1", "chair2", "chair4", "desk1", "desk1", "desk3", "sofa", "Chair_025"]


def load(name):
    pts = to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", name + ".pcd")))
    if name == "Chair_025":
        pts[:, :3] *= np.float32(0.01)
    return pts


ctx = api.Context(0)
feats = []
for m in MODELS + ["mcloud"]:
    c = api.Cloud(ctx, load(m))
    c.normals(0.05)
    feats.append(c.fpfh(0.10))
    c.free()
fa, fb = np.concatenate(feats[:-1]), feats[-1]
ms, st = api.match_raw(ctx, fa, fb, 5, reps=5)
print("keep=%s tc=%s : %d x %d  %.1f us  redo rows %d  splits %d" % (os.environ.get("RTR_MATCH_KEEP", "auto"), os.environ.get("RTR_MATCH_TC", "auto"),
                                                                  len(fa), len(fb), 1e3 * ms, st["redo_rows"], st["splits"]))
