#!/usr/bin/env python3
"""Per-stage device time (library profiling marks) of one registration per repo model against mcloud, each alone."""
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from realtime_robot_b200 import api  # noqa: E402
from realtime_robot_b200.pcd import read_pcd_xyz, to_xyz1  # noqa: E402

names = sys.argv[1:] or ["chair1", "chair4", "sofa"]
ctx = api.Context(0)
s = to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", "mcloud.pcd")))
p = api.default_register_params()
for name in names:
    m = to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", name + ".pcd")))
    if name == "Chair_025":
        m[:, :3] *= 0.01
    for _ in range(3):
        api.register_host(ctx, m, s, p)
    ctx.profile_begin()
    r = api.register_host(ctx, m, s, p)
    pr = ctx.profile_end()
    tot = sum(v[1] for v in pr.values())
    print(name, len(m), "points; serialised device ms", round(tot, 3), "inliers", r.inliers)
    for k, v in sorted(pr.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("TOP", "14"))]:
        print("   %-22s x%-3d %.3f ms" % (k, v[0], v[1]))
