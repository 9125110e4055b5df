#!/usr/bin/env python3
"""Multi-GPU check of the record exchange (csrc/comm.cu), run under torchrun: synthetic records with rank- and step-dependent
contents are all-gathered 200 times through rtr_allgather_results — once with the peer-memory exchange (RTR_COMM_P2P=1), once with
the default ncclAllGather — every gathered byte is checked on every rank, and the mean latency of both is printed by
rank 0.  Then a batch per rank with rtr_comm_gather_batches (rank r registers chair1 + chair2 against mcloud with r-dependent
hypothesis counts) is compared with the records every rank computes for every other rank's settings.
usage: python -m torch.distributed.run --nproc-per-node N tools/check_comm.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as td  # noqa: E402
from realtime_robot_b200 import _lib, api, dist  # noqa: E402
from realtime_robot_b200.params import PoseResult, default_register_params  # noqa: E402
from realtime_robot_b200.pcd import read_pcd_xyz, to_xyz1  # noqa: E402

rank, local_rank, world = dist.env_world()
torch.cuda.set_device(local_rank)
td.init_process_group("gloo", rank=rank, world_size=world)        # rendezvous only: the exchange under test is the library's own
out = {"world": world}
load = lambda n: to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", n + ".pcd")))


def synthetic(r, step, n):
    recs = []
    for k in range(n):
        x = PoseResult()
        x.model_id = r * 100 + k; x.hypothesis = step * 1000 + r; x.inliers = k; x.fitness = float(r) + 0.001 * step
        for i in range(16):
            x.pose[i] = r + i + 0.5 * step
        recs.append(x)
    return recs


for mode in ("p2p", "nccl"):
    os.environ["RTR_COMM_P2P"] = "1" if mode == "p2p" else "0"
    ctx = api.Context(local_rank)
    dist.comm_init(ctx, rank, world)
    ok = True
    for n in (1, 8, 64):
        for step in range(20):
            got = dist.allgather_results(ctx, synthetic(rank, step, n), world)
            want = [x for r in range(world) for x in synthetic(r, step, n)]
            ok &= len(got) == len(want) and all(bytes(a) == bytes(b) for a, b in zip(got, want))
    mine = synthetic(rank, 7, 8)
    td.barrier()
    t0 = time.perf_counter()
    for _ in range(200):
        dist.allgather_results(ctx, mine, world)
    us = 1e6 * (time.perf_counter() - t0) / 200
    # a batch per rank, gathered in-stream
    p = default_register_params()
    p.ransac.max_iterations = 3000 + 500 * rank
    models, scene = [load("chair1"), load("chair2")], load("mcloud")
    dist.gather_batches(ctx, True, rank * 2)
    api.register_many_host(ctx, models, scene, p)
    allr, n_all = dist.gathered_results(ctx, 64)
    dist.gather_batches(ctx, False)
    batch_ok = n_all == 2 * world
    for r in range(world):
        q = default_register_params()
        q.ransac.max_iterations = 3000 + 500 * r
        want = api.register_many_host(ctx, models, scene, q)
        for k in range(2):
            g = allr[r * 2 + k]
            batch_ok &= g.model_id == r * 2 + k and (g.hypothesis, g.inliers, np.float32(g.fitness).tobytes(), bytes(g.pose)) == \
                (want[k].hypothesis, want[k].inliers, np.float32(want[k].fitness).tobytes(), bytes(want[k].pose))
    flags = torch.tensor([int(ok), int(batch_ok)], dtype=torch.int32)
    td.all_reduce(flags, op=td.ReduceOp.MIN)
    out[mode] = {"records_ok_on_every_rank": bool(flags[0]), "gathered_batch_ok_on_every_rank": bool(flags[1]), "allgather_results_us": round(us, 2)}
    _lib.lib().rtr_comm_destroy(ctx._h)
    td.barrier()
if rank == 0:
    print(json.dumps(out))
