#!/usr/bin/env python3
"""One bench step (rtr_register_many: mcloud vs the 8 repo models) for profilers: warm-up steps, then `--steps` steps between
cudaProfilerStart / cudaProfilerStop (run ncu with --profile-from-start off).  Prints the unprofiled step time first."""
import argparse
import ctypes
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


ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--models", default=",".join(MODELS))
ap.add_argument("--marks", action="store_true", help="print the per-operation device times of one step (profiling marks)")
ap.add_argument("--prepared", action="store_true", help="the online half instead: models prepared once (rtr_cloud_prepare), a step = "
                "upload of the scan + rtr_register_prepared")
args = ap.parse_args()
names = args.models.split(",")
ctx = api.Context(0)
scene = api.Cloud(ctx, load("mcloud"))
models = [api.Cloud(ctx, load(m)) for m in names]
p = api.default_register_params()
if args.prepared:
    scene_h = load("mcloud")
    for m in models:
        m.prepare(p)

    def step():
        cs = api.Cloud(ctx, scene_h)
        r = api.register_prepared(models, cs, p)
        cs.free()
        return r
    for _ in range(args.warmup):
        step()
    ts = []
    for _ in range(10):
        ctx.sync(); t0 = time.perf_counter(); step(); ts.append(1e3 * (time.perf_counter() - t0))
    print("online step %.3f ms (min of 10, wall)" % min(ts), flush=True)
    if args.marks:
        ctx.profile_begin(); step(); pr = ctx.profile_end()
        tot = sum(v[1] for v in pr.values())
        for k, v in sorted(pr.items(), key=lambda kv: -kv[1][1]):
            print("  %-22s x%-3d %8.1f us  %5.1f %%" % (k, v[0], 1e3 * v[1], 100 * v[1] / tot))
        print("  serialised %.3f ms" % tot)
    rt = ctypes.CDLL("libcudart.so")
    rt.cudaProfilerStart()
    for _ in range(args.steps):
        step()
    ctx.sync()
    rt.cudaProfilerStop()
    sys.exit(0)
for _ in range(args.warmup):
    api.register_many(models, scene, p)
ts = []
for _ in range(10):
    ctx.sync(); t0 = time.perf_counter(); api.register_many(models, scene, p); ts.append(1e3 * (time.perf_counter() - t0))
print("step %.3f ms (min of 10, wall), launches/step %d" % (min(ts), 0), flush=True)
if args.marks:
    ctx.profile_begin(); api.register_many(models, scene, p); pr = ctx.profile_end()
    tot = sum(v[1] for v in pr.values())
    for k, v in sorted(pr.items(), key=lambda kv: -kv[1][1]):
        print("  %-22s x%-3d %8.1f us  %5.1f %%" % (k, v[0], 1e3 * v[1], 100 * v[1] / tot))
    print("  serialised %.3f ms" % tot)
# host-side issue time of the two entry points (begin = everything queued, end = wait + records)
hosts = [load(m) for m in names]
scene_h = load("mcloud")
import torch  # noqa: E402
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
hosts_p, scene_p = [pin(h) for h in hosts], pin(scene_h)
for label, begin, end in (("resident", lambda: api.register_many_begin(models, scene, p), lambda: api.register_many_end(ctx)),
                          ("host    ", lambda: api.register_many_host_begin(ctx, hosts_p, scene_p, p), lambda: api.register_many_end(ctx))):
    tb, tt = [], []
    for _ in range(12):
        ctx.sync(); t0 = time.perf_counter(); begin(); t1 = time.perf_counter(); end(); t2 = time.perf_counter()
        tb.append(1e3 * (t1 - t0)); tt.append(1e3 * (t2 - t0))
    print("%s: begin (enqueue) %.3f ms, begin+end %.3f ms (min of 12)" % (label, min(tb), min(tt)), flush=True)
rt = ctypes.CDLL("libcudart.so")
rt.cudaProfilerStart()
for _ in range(args.steps):
    api.register_many(models, scene, p)
ctx.sync()
rt.cudaProfilerStop()
