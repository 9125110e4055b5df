#!/usr/bin/env python3
"""BASELINE.json configs[4]: prerejective RANSAC sweep 1e4..1e7 hypotheses, hypothesis-sharded over the ranks of a
torchrun job (one process per GPU, NCCL): rank r evaluates the contiguous range dist.shard_hypotheses gives it, the
128-byte records are all-gathered once, every rank takes the arg-min over (fitness, hypothesis).  Prints one JSON line
per sweep point on rank 0: device time = max over ranks (CUDA events on the library's stream), plus whether the sharded
winner equals the unsharded one.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_ransac_sharded.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from realtime_robot_b200 import api, dist  # noqa: E402
from realtime_robot_b200.pcd import read_pcd_xyz, to_xyz1  # noqa: E402


def main():
    rank, local_rank, world = dist.env_world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    ctx = api.Context(local_rank)
    m = to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", "chair1.pcd")))
    s = to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", "mcloud.pcd")))
    cm, cs = api.Cloud(ctx, m), api.Cloud(ctx, s)
    p = api.default_register_params()
    for c in (cm, cs):
        c.normals(p.normal_radius); c.fpfh(p.fpfh_radius)
    cm.match_features(cs, 5)
    H = 10_000
    while H <= 10_000_000:
        p.ransac.max_iterations = H
        p.ransac.hypothesis_begin, p.ransac.hypothesis_end = 0, 0
        full = api.ransac_prerejective(cm, cs, p.ransac) if rank == 0 else None
        b, e = dist.shard_hypotheses(H, rank, world)
        p.ransac.hypothesis_begin, p.ransac.hypothesis_end = b, e
        api.ransac_prerejective(cm, cs, p.ransac)                      # warm-up
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()
        ctx.record(4); r = api.ransac_prerejective(cm, cs, p.ransac); ctx.record(5)
        ms = torch.tensor([ctx.elapsed_ms(4, 5)], device=dev)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        recs = dist.all_gather_records([r], 1, device=dev)
        t1.record(); torch.cuda.synchronize()
        if world > 1:
            td.all_reduce(ms, op=td.ReduceOp.MAX)
        best = dist.select_best_hypothesis(recs)
        if rank == 0:
            same = best is not None and full is not None and (best.hypothesis, best.inliers) == (full.hypothesis, full.inliers) and \
                np.array_equal(np.asarray(best.pose), np.asarray(full.pose))
            print(json.dumps({"config": "configs[4] prerejective RANSAC sweep, hypothesis-sharded", "n_gpus": world, "hypotheses": H,
                              "ms_max_over_ranks": float(ms.item()), "allgather_ms": t0.elapsed_time(t1),
                              "hypotheses_per_s": H / (float(ms.item()) * 1e-3), "winner": int(best.hypothesis) if best else -1,
                              "inliers": int(best.inliers) if best else 0, "equals_unsharded_winner": bool(same)}), flush=True)
        H *= 10
    if world > 1:
        td.barrier()
        td.destroy_process_group()


if __name__ == "__main__":
    main()
