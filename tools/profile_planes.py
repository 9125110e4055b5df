#!/usr/bin/env python3
"""Device time per stage (library profiling marks) and wall time of rtr_plane_areas on repo clouds."""
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from realtime_robot_b200 import api  # noqa: E402
from realtime_robot_b200.pcd import read_pcd_xyz, to_xyz1  # noqa: E402

ctx = api.Context(0)
for name in sys.argv[1:] or ["chair1", "T0_m8111", "chair4"]:
    c = api.Cloud(ctx, to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", name + ".pcd"))))
    for _ in range(3):
        s = api.plane_areas(c)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(10):
        s = api.plane_areas(c)
    wall = (time.perf_counter() - t0) / 10
    ctx.profile_begin()
    s = api.plane_areas(c)
    pr = ctx.profile_end()
    tot = sum(v[1] for v in pr.values())
    print(f"{name}: {c.n} points, {len(s)} planes, wall {wall * 1e3:.3f} ms, marked device {tot:.3f} ms, launches {sum(v[0] for v in pr.values())}")
    for k, v in sorted(pr.items(), key=lambda kv: -kv[1][1])[:8]:
        print("   %-22s x%-3d %.3f ms" % (k, v[0], v[1]))
