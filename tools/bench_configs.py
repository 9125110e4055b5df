#!/usr/bin/env python3
"""Throughput of the other configurations BASELINE.json names (they are parity-test / scale cases, not the bench line):
  configs[3]  synthetic 4M-point scene: normals (r=.05) + FPFH (r=.08, all points) ; feature matching 262144 x 65536 x 33
  configs[4]  prerejective RANSAC sweep 1e4..1e7 hypotheses on the config-1 clouds
Prints one JSON object per section.  usage: bench_configs.py [--points 4000000] [--match-m 262144] [--match-n 65536]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from realtime_robot_b200 import api, synth  # noqa: E402
from realtime_robot_b200.pcd import read_pcd_xyz, to_xyz1  # noqa: E402

HBM = 6546.6
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def timed(ctx, fn, reps=3):
    fn()
    ms = []
    for _ in range(reps):
        ctx.sync(); ctx.record(4); fn(); ctx.record(5)
        ms.append(ctx.elapsed_ms(4, 5))
    return min(ms), ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=4_000_000)
    ap.add_argument("--match-m", type=int, default=262144)
    ap.add_argument("--match-n", type=int, default=65536)
    ap.add_argument("--max-hyp", type=float, default=1e7)
    ap.add_argument("--skip", default="")
    args = ap.parse_args()
    ctx = api.Context(0)

    if "scene" not in args.skip:
        n = args.points
        side = max(4.0, (-12.0 + np.sqrt(144.0 + 4.0 * (n / 4700.0 * 0.8))) / 2.0)
        t0 = time.time()
        scene = synth.sample_rects(synth.room_rects((side, side, 3.0), n_boxes=max(4, int(side)), seed=synth.BASE_SEED), n, synth.BASE_SEED + 7)
        print(f"# generated {n} points, room {side:.1f} m ({time.time() - t0:.1f}s)", file=sys.stderr, flush=True)
        c = api.Cloud(ctx, scene)

        def normals():
            c.reset(); _lib.rtr_normals(c._h, 0.05, None)
        from realtime_robot_b200 import _lib as L
        _lib = L.lib()
        ms_n, _ = timed(ctx, normals)
        cnt5, _, _ = c.radius_neighbors(0.05, counts_only=True)
        k5 = int(cnt5.sum())
        ctx.profile_begin(); normals(); pr = ctx.profile_end()
        kms = pr["normals"][1]
        alg = 16 * (n + k5) + 16 * n
        out = {"config": "configs[3] normals", "points": n, "radius": 0.05, "sum_neighbours": k5, "mean_neighbours": k5 / n,
               "ms_incl_grid_build": ms_n, "points_per_s": n / (ms_n * 1e-3), "kernel_ms": kms,
               "roofline": {"kernel": "k_normals", "bound": "hbm", "achieved": alg / (kms * 1e-3) / 1e9, "peak": HBM, "unit": "GB/s",
                            "frac": alg / (kms * 1e-3) / 1e9 / HBM, "algorithmic_bytes": alg},
               "kernel_ms_all": {k: round(v[1], 3) for k, v in sorted(pr.items(), key=lambda kv: -kv[1][1])[:6]}}
        print(json.dumps(out), flush=True)

        def fpfh():
            _lib.rtr_cloud_reset(c._h); _lib.rtr_normals(c._h, 0.05, None); _lib.rtr_fpfh(c._h, 0.08, None)
        ms_f, _ = timed(ctx, fpfh, reps=2)
        cnt8, _, _ = c.radius_neighbors(0.08, counts_only=True)
        k8 = int(cnt8.sum())
        ctx.profile_begin(); fpfh(); pr = ctx.profile_end()
        alg_s = 16 * (n + k8) + 16 * k8 + 132 * n
        alg_w = 16 * (n + k8) + 132 * k8 + 132 * n
        out = {"config": "configs[3] FPFH (all points are queries)", "points": n, "radius": 0.08, "sum_neighbours": k8,
               "mean_neighbours": k8 / n, "ms_normals_plus_fpfh": ms_f, "queries_per_s": n / (ms_f * 1e-3),
               "neighbour_entries_per_s": k8 / (ms_f * 1e-3),
               "roofline_spfh": {"kernel": "k_spfh", "kernel_ms": pr["fpfh.spfh"][1], "achieved": alg_s / (pr["fpfh.spfh"][1] * 1e-3) / 1e9,
                                 "frac": alg_s / (pr["fpfh.spfh"][1] * 1e-3) / 1e9 / HBM, "algorithmic_bytes": alg_s, "peak": HBM, "unit": "GB/s"},
               "roofline_weight": {"kernel": "k_fpfh_weight", "kernel_ms": pr["fpfh.weight"][1], "achieved": alg_w / (pr["fpfh.weight"][1] * 1e-3) / 1e9,
                                   "frac": alg_w / (pr["fpfh.weight"][1] * 1e-3) / 1e9 / HBM, "algorithmic_bytes": alg_w, "peak": HBM, "unit": "GB/s"},
               "kernel_ms_all": {k: round(v[1], 3) for k, v in sorted(pr.items(), key=lambda kv: -kv[1][1])[:8]}}
        print(json.dumps(out), flush=True)
        # features of the scene, reused as the matching workload (real FPFH statistics, not random numbers)
        feats = c.fpfh(0.08)
        # SURVEY 8(d) configs[3], query form: 100 000 keypoints (one point per occupied voxel of a coarse lattice, then a
        # seeded subsample) on the full 4 M-point search surface, r = 0.08 -> ~10 M neighbour entries in the weighting pass
        vox = np.floor(scene[:, :3] / 0.06).astype(np.int64)
        _, first = np.unique(vox[:, 0] * 1_000_003 + vox[:, 1] * 1009 + vox[:, 2], return_index=True)
        rng_q = np.random.default_rng(synth.BASE_SEED + 3)
        qidx = np.sort(rng_q.choice(first, min(100_000, len(first)), replace=False)).astype(np.int32)
        sub = c.fpfh_at(0.08, qidx)
        same = bool(np.array_equal(sub.view(np.uint32), feats[qidx].view(np.uint32)))
        ms_q = []
        for _ in range(3):
            ctx.sync(); ctx.record(4); c.fpfh_at(0.08, qidx); ctx.record(5)
            ms_q.append(ctx.elapsed_ms(4, 5))
        ctx.profile_begin(); c.fpfh_at(0.08, qidx); pr = ctx.profile_end()
        kq = int(cnt8[qidx].sum())
        alg_wq = 16 * (len(qidx) + kq) + 132 * kq + 132 * len(qidx)
        print(json.dumps({"config": "configs[3] FPFH at 100k keypoints on the 4M-point surface (rtr_fpfh_at, host query list in, rows out)",
                          "queries": int(len(qidx)), "radius": 0.08, "neighbour_entries": kq, "rows_equal_full_fpfh_bitwise": same,
                          "ms_end_to_end": min(ms_q), "queries_per_s": len(qidx) / (min(ms_q) * 1e-3),
                          "roofline_weight": {"kernel": "k_fpfh_weight", "kernel_ms": pr["fpfh.weight"][1], "algorithmic_bytes": alg_wq,
                                              "achieved": alg_wq / (pr["fpfh.weight"][1] * 1e-3) / 1e9, "peak": HBM, "unit": "GB/s",
                                              "frac": alg_wq / (pr["fpfh.weight"][1] * 1e-3) / 1e9 / HBM},
                          "kernel_ms_all": {k: round(v[1], 3) for k, v in sorted(pr.items(), key=lambda kv: -kv[1][1])[:8]}}), flush=True)
        c.free()
    else:
        feats = None

    if "match" not in args.skip:
        M, N = args.match_m, args.match_n
        rng = np.random.default_rng(synth.BASE_SEED)
        if feats is None:
            feats = rng.random((max(M, N), 33)).astype(np.float32) * 30
        fa = np.ascontiguousarray(feats[rng.choice(len(feats), M, replace=len(feats) < M)])
        fb = np.ascontiguousarray(feats[rng.choice(len(feats), N, replace=len(feats) < N)])
        ms, stats = api.match_raw(ctx, fa, fb, 5, reps=2)
        flop = 2.0 * M * N * 33
        out = {"config": "configs[3] descriptor matching", "M": M, "N": N, "k": 5, "ms": ms, "pairs_per_s": M * N / (ms * 1e-3),
               "algorithmic_tflops": flop / (ms * 1e-3) / 1e12,
               "rows_redone_exactly": stats["redo_rows"], "splits": stats["splits"], "observed_err_over_norms": stats["observed_err_over_norms"], "idx0": stats["idx"][0].tolist(), "dist0": stats["dist"][0].tolist()}
        print(json.dumps(out), flush=True)

    if "ransac" not in args.skip:
        m = to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", "chair1.pcd")))
        s = to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", "mcloud.pcd")))
        cm, cs = api.Cloud(ctx, m), api.Cloud(ctx, s)
        p = api.default_register_params()
        for c in (cm, cs):
            c.normals(p.normal_radius); c.fpfh(p.fpfh_radius)
        cm.match_features(cs, 5)
        H = 1e4
        while H <= args.max_hyp:
            p.ransac.max_iterations = int(H)
            r = api.ransac_prerejective(cm, cs, p.ransac)
            ctx.sync(); ctx.record(4); r = api.ransac_prerejective(cm, cs, p.ransac); ctx.record(5)
            ms = ctx.elapsed_ms(4, 5)
            print(json.dumps({"config": "configs[4] prerejective RANSAC sweep", "hypotheses": int(H), "ms": ms, "hypotheses_per_s": H / (ms * 1e-3),
                              "survivors": int(r.evaluated), "prerejection_rate": 1.0 - r.evaluated / H, "winner": int(r.hypothesis),
                              "inliers": int(r.inliers), "fitness": float(r.fitness),
                              "survivor_evals_per_s": r.evaluated / (ms * 1e-3)}), flush=True)
            H *= 10


if __name__ == "__main__":
    main()
