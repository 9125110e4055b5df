#!/usr/bin/env python3
"""ICP @ 1 M points (BASELINE.json configs[2]) on its own: both directions, per-kernel times, result bytes.
Development A/B aid (the RTR_* environment switches of DESIGN.md select variants); the graded
numbers are bench.py's icp_1m section.  usage: bench_icp.py [--model 100000] [--scan 1000000] [--iters 50]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from realtime_robot_b200 import api, synth  # noqa: E402
from realtime_robot_b200.params import default_register_params  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", type=int, default=100_000)
    ap.add_argument("--scan", type=int, default=1_000_000)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    ctx = api.Context(0)
    model, scan, gt = synth.icp_config(args.model, args.scan)
    cm, cs = api.Cloud(ctx, model), api.Cloud(ctx, scan)
    q = default_register_params()
    q.icp.max_iterations = args.iters
    q.icp.force_iterations = 1
    out = {"env": {k: v for k, v in os.environ.items() if k.startswith("RTR_")}}
    for label, a, b, cap in (("model_to_scan", cm, cs, 0.0), ("scan_to_model", cs, cm, 0.05), ("scan_to_model_uncapped", cs, cm, 0.0)):
        q.icp.max_correspondence_distance = cap
        for _ in range(2):
            a.reset(); b.reset(); api.icp(a, b, q.icp, None)
        ms = []
        for _ in range(args.reps):
            a.reset(); b.reset(); ctx.sync()
            ctx.record(2); res = api.icp(a, b, q.icp, None); ctx.record(3)
            ms.append(ctx.elapsed_ms(2, 3))
        a.reset(); b.reset()
        ctx.profile_begin(); api.icp(a, b, q.icp, None); pr = ctx.profile_end()
        out[label] = {"ms_per_icp": round(min(ms), 4), "iters_per_s": round(args.iters / (min(ms) * 1e-3), 1),
                      "corr_us_per_launch": round(1e3 * pr["icp.corr"][1] / pr["icp.corr"][0], 2),
                      "fitness_ms": round(pr["icp.fitness"][1], 4), "inliers": int(res.inliers), "fitness": float(res.fitness),
                      "pose_sum": float(np.abs(res.matrix()).sum()), "result_hex": bytes(res).hex()[:64]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
