#!/usr/bin/env python3
"""Golden numbers mined from the reference's own output artefacts (run in the build container only).

RealTimeRobot/transformed_cloud1.pcd (== transformed_cloud3.pcd) is T0_m8111.pcd pushed through the matrix an earlier
version of main() computed (RealTimeRobot.cpp:104-109 writes `cloud` transformed by Ransac()'s result).  That matrix is
get_Distance's key_transform for ONE (model keypoint, scan keypoint) pair: a yaw step about the model keypoint, the
translation k_model - k_scan and the scale z_model / z_scan (matching.h:204-217).  A least-squares affine fit between
the two files therefore exposes four numbers that depend only on the two Harris corners PCL 1.8.0 produced:
    scale = z_m / z_s,   t_z = scale * (z_m - z_s),   t_x = x_m - cos(yaw) x_s + sin(yaw) y_s,   t_y' (scaled row)
They are written to tests/golden/reference_artifacts.json and pin the keypoint chain of the oracle and the CUDA path."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from realtime_robot_b200.pcd import read_pcd_xyz  # noqa: E402

REF = "/root/reference/RealTimeRobot/"
src = read_pcd_xyz(REF + "T0_m8111.pcd").astype(np.float64)
out = {}
for name in ("transformed_cloud1.pcd", "transformed_cloud3.pcd"):
    dst = read_pcd_xyz(REF + name).astype(np.float64)
    A = np.c_[src, np.ones(len(src))]
    M = np.linalg.lstsq(A, dst, rcond=None)[0].T
    resid = float(np.abs(A @ M.T - dst).max())
    yaw = float(np.degrees(np.arctan2(-M[0, 1], M[0, 0])))
    out[name] = {"affine_3x4": M.tolist(), "max_residual": resid, "yaw_deg": yaw, "scale_z": float(M[2, 2]),
                 "t": [float(M[0, 3]), float(M[1, 3]), float(M[2, 3])],
                 "row_scale": [float(np.hypot(M[r, 0], M[r, 1])) for r in range(2)],
                 "model": "chair1.pcd", "scan": "T0_m8111.pcd"}
# Two more artefacts of the same kind (round 2): mcloudYasuo.pcd and transformed_cloud2.pcd are ALSO T0_m8111.pcd pushed through a
# get_Distance transform — yaw 70 and 160 degrees, one xy row scaled by z_m / z_s = 1.070555 (the other xy row is the one scaled
# in transformed_cloud1: the scaling moved between code versions).  No Harris corner pair of ANY cloud the reference ships gives
# that z ratio against T0_m8111's corners (searched: every .pcd, refined corners, |ratio - 1.070555| < 2e-4), and the sign of t_z
# contradicts t_z = s (z_m - z_s) for positive corner heights: the model cloud of that run is not among the shipped files.  So
# they pin the FORM of the sweep transform (yaw quantised to the 10 degree step of matching.h:143, one similarity scale shared
# by the z row and one xy row), not another corner pair.  70761_c.pcd = 70761.pcd + (1, 0.5, 0) and mcloud1.pcd = chair1.pcd are
# plain copies (weak fixtures).
for name in ("mcloudYasuo.pcd", "transformed_cloud2.pcd"):
    d2 = read_pcd_xyz(REF + name).astype(np.float64)
    A = np.c_[src, np.ones(len(src))]
    M = np.linalg.lstsq(A, d2, rcond=None)[0].T
    rows = [float(np.hypot(M[r, 0], M[r, 1])) for r in range(2)]
    unscaled = int(np.argmin([abs(r - 1.0) for r in rows]))
    yaw = float(np.degrees(np.arctan2(M[1, 0] / rows[1], M[0, 0] / rows[0])))
    out[name] = {"affine_3x4": M.tolist(), "max_residual": float(np.abs(A @ M.T - d2).max()), "yaw_deg": yaw % 360.0, "scale_z": float(M[2, 2]),
                 "t": [float(M[0, 3]), float(M[1, 3]), float(M[2, 3])], "row_scale": rows, "unscaled_row": unscaled,
                 "scan": "T0_m8111.pcd", "model": "not among the shipped clouds (see tools/fit_reference_artifacts.py)"}
dst = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_artifacts.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out["transformed_cloud1.pcd"], indent=1))
