#!/usr/bin/env python3
"""Descriptor matching at BASELINE.json configs[3] size on its own: 262 144 x 65 536 x 33 on real FPFH rows of the 4 M-point
synthetic scene, k = 5.  Prints the device time of the whole search, the per-kernel times of one more (profiled) run and the
rows the certificate sent to the exact kernel.  Development A/B aid (RTR_MATCH_* switches); the graded number is bench.py's
matching_262k_x_65k section.  usage: bench_match_scale.py [--points 4000000] [--m 262144] [--n 65536] [--check 2048]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from realtime_robot_b200 import _lib, api, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=4_000_000)
    ap.add_argument("--m", type=int, default=262144)
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--check", type=int, default=512, help="rows compared with a numpy fp64 brute-force search")
    ap.add_argument("--variants", default="", help="';'-separated RTR_* settings to time one after the other, e.g. 'RTR_MATCH_KEEP=8,RTR_MATCH_SHARE=0;RTR_MATCH_KEEP=16'")
    args = ap.parse_args()
    ctx = api.Context(0)
    L = _lib.lib()
    side = max(4.0, (-12.0 + np.sqrt(144.0 + 4.0 * (args.points / 4700.0 * 0.8))) / 2.0)
    scene = synth.sample_rects(synth.room_rects((side, side, 3.0), n_boxes=max(4, int(side)), seed=synth.BASE_SEED), args.points, synth.BASE_SEED + 7)
    c4 = api.Cloud(ctx, scene)
    L.rtr_normals(c4._h, 0.05, None)
    feats = c4.fpfh(0.08)
    c4.free()
    rng = np.random.default_rng(synth.BASE_SEED)
    fa = np.ascontiguousarray(feats[rng.choice(len(feats), args.m, replace=False)])
    fb = np.ascontiguousarray(feats[rng.choice(len(feats), args.n, replace=False)])
    del feats
    for variant in [v for v in args.variants.split(";") if v.strip()]:
        saved = {k: os.environ.get(k) for k in [kv.split("=")[0] for kv in variant.split(",")]}
        for kv in variant.split(","):
            k, v = kv.split("=")
            os.environ[k] = v
        api.match_raw(ctx, fa, fb, 5)
        ms, st = api.match_raw(ctx, fa, fb, 5, reps=args.reps)
        ctx.profile_begin(); api.match_raw(ctx, fa, fb, 5); pr = ctx.profile_end()
        print(json.dumps({"variant": variant, "ms": round(ms, 3), "redo_rows": st["redo_rows"], "err_over_norms": st["observed_err_over_norms"],
                          "kernels_ms": {k: round(v[1], 3) for k, v in sorted(pr.items(), key=lambda kv: -kv[1][1]) if k.startswith("match")}}), flush=True)
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    api.match_raw(ctx, fa, fb, 5)
    ms, st = api.match_raw(ctx, fa, fb, 5, reps=args.reps)
    ctx.profile_begin()
    api.match_raw(ctx, fa, fb, 5)
    pr = ctx.profile_end()
    out = {"env": {k: v for k, v in os.environ.items() if k.startswith("RTR_")}, "M": args.m, "N": args.n, "ms": round(ms, 3),
           "algorithmic_tflops": round(2.0 * args.m * args.n * 33 / (ms * 1e-3) / 1e12, 1), "redo_rows": st["redo_rows"], "splits": st["splits"],
           "err_over_norms": st["observed_err_over_norms"],
           "kernels_ms": {k: round(v[1], 3) for k, v in sorted(pr.items(), key=lambda kv: -kv[1][1]) if k.startswith("match")}}
    if args.check > 0:
        rows = rng.choice(args.m, args.check, replace=False)
        a = fa[rows].astype(np.float64)
        b = fb.astype(np.float64)
        bad = 0
        for lo in range(0, len(rows), 32):
            d = np.zeros((len(a[lo:lo + 32]), len(b)))
            for e in range(33):                  # the oracle's sequential fp64 sum over the 33 bins
                d += (a[lo:lo + 32, e, None] - b[None, :, e]) ** 2
            d = d.astype(np.float32)            # the oracle compares the float-rounded distances, ties -> lowest index
            order = np.lexsort((np.broadcast_to(np.arange(d.shape[1]), d.shape), d), axis=1)[:, :5]
            bad += int((order != st["idx"][rows[lo:lo + 32]]).any(axis=1).sum())
        out["rows_checked"] = int(len(rows))
        out["rows_different_from_numpy_fp64"] = bad
    print(json.dumps(out))


if __name__ == "__main__":
    main()
