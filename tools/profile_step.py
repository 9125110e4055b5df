#!/usr/bin/env python3
"""The bench step (8 registrations vs mcloud on 8 contexts, queued by one host thread) with the profiling marks of every
context on: per-context span and per-stage durations UNDER CONTENTION, beside the same registration run alone."""
import os
import sys
import time

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


ctxs = [api.Context(0) for _ in MODELS]
scene_h = load("mcloud")
ms = [api.Cloud(c, load(m)) for c, m in zip(ctxs, MODELS)]
ss = [api.Cloud(c, scene_h) for c in ctxs]
p = api.default_register_params()
order = sorted(range(len(MODELS)), key=lambda i: -ms[i].n)


def step(profile):
    for c in ctxs:
        c.sync()
    if profile:
        for c in ctxs:
            c.profile_begin()
    t0 = time.perf_counter()
    for i in order:
        ms[i].reset(); ss[i].reset()
        api.register_begin(ms[i], ss[i], p)
    t_issue = time.perf_counter() - t0
    for c in ctxs:
        api.register_end(c)
    wall = time.perf_counter() - t0
    prof = [c.profile_end() for c in ctxs] if profile else None
    return wall, t_issue, prof


for _ in range(4):
    step(False)
w = [step(False) for _ in range(10)]
print("step wall %.3f ms (issue %.3f ms), unprofiled" % (1e3 * np.mean([x[0] for x in w]), 1e3 * np.mean([x[1] for x in w])))
wall, t_issue, prof = step(True)
print("profiled step wall %.3f ms" % (1e3 * wall))
for i in order:
    tot = sum(v[1] for v in prof[i].values())
    top = sorted(prof[i].items(), key=lambda kv: -kv[1][1])[:6]
    print("%-10s n=%-6d span %.3f ms : %s" % (MODELS[i], ms[i].n, tot, ", ".join("%s %.2f" % (k, v[1]) for k, v in top)))
