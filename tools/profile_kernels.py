#!/usr/bin/env python3
"""Workload for `ncu --set full` captures (no torch import: fast start-up).  Two chair4 -> mcloud registrations
(the densest repo model: 18779 points), then a 5-iteration ICP of the 1M-point scan onto the 100k-point model."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from realtime_robot_b200 import api, synth  # noqa: E402
from realtime_robot_b200.pcd import read_pcd_xyz, to_xyz1  # noqa: E402

ctx = api.Context(0)
m = to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", "chair4.pcd")))
s = to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", "mcloud.pcd")))
p = api.default_register_params()
for _ in range(2):
    r = api.register_host(ctx, m, s, p)
print("registration", r.hypothesis, r.inliers, r.fitness)
model, scan, gt = synth.icp_config(100_000, 1_000_000)
cm, cs = api.Cloud(ctx, model), api.Cloud(ctx, scan)
p.icp.max_iterations = 5
p.icp.force_iterations = 1
p.icp.max_correspondence_distance = 0.05
r = api.icp(cs, cm, p.icp)
print("icp scan->model", r.iterations, r.fitness)
p.icp.max_correspondence_distance = 0.0
r = api.icp(cm, cs, p.icp)
print("icp model->scan", r.iterations, r.fitness)
