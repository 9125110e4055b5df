"""Multi-GPU plumbing: one process per GPU, torch.distributed for the rendezvous and for ONE small
all-gather of fixed 128-byte pose records per registration batch.  The path shards by candidate model cloud
(rank r takes models m with m mod W == r) and by RANSAC hypothesis range; there is no data-path collective.
"""
import ctypes as C
import os

import numpy as np

from .params import PoseResult

RECORD_BYTES = C.sizeof(PoseResult)


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_models(n_models: int, rank: int, world: int):
    """Model ids of this rank: round-robin, so the result is independent of how many models follow."""
    return list(range(rank, n_models, world))


def shard_hypotheses(max_iterations: int, rank: int, world: int):
    """Contiguous hypothesis range [begin, end) of this rank (hypothesis h depends only on (seed, h))."""
    per = (max_iterations + world - 1) // world
    b = min(rank * per, max_iterations)
    return b, min(b + per, max_iterations)


def records_to_bytes(records) -> np.ndarray:
    buf = np.zeros((len(records), RECORD_BYTES), dtype=np.uint8)
    for i, r in enumerate(records):
        buf[i] = np.frombuffer(bytes(r), dtype=np.uint8)
    return buf


def bytes_to_records(buf: np.ndarray):
    out = []
    for row in np.ascontiguousarray(buf, dtype=np.uint8).reshape(-1, RECORD_BYTES):
        out.append(PoseResult.from_buffer_copy(row.tobytes()))
    return out


def all_gather_records(records, per_rank: int, device=None):
    """Gather `per_rank` records from every rank (ranks with fewer pad with hypothesis = -2 records, dropped on return).
    One torch.distributed.all_gather_into_tensor of world * per_rank * 128 bytes; NCCL when `device` is a cuda device."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    pad = PoseResult()
    pad.hypothesis = -2
    pad.model_id = -1
    recs = list(records) + [pad] * (per_rank - len(records))
    local = torch.from_numpy(records_to_bytes(recs).reshape(-1))
    if world == 1:
        gathered = local
    else:
        if device is not None:
            local = local.to(device)
        gathered = torch.empty(world * local.numel(), dtype=torch.uint8, device=local.device)
        dist.all_gather_into_tensor(gathered, local)
        gathered = gathered.cpu()
    out = bytes_to_records(gathered.numpy())
    return [r for r in out if not (r.hypothesis == -2 and r.model_id == -1)]


def select_best_hypothesis(records):
    """Deterministic arg-min over (fitness, hypothesis id) among accepted shards: identical on every rank and for
    every world size (the sequential PCL rule 'error < lowest_error' keeps the first lowest)."""
    ok = [r for r in records if r.hypothesis >= 0 and r.converged]
    if not ok:
        return None
    return min(ok, key=lambda r: (float(r.fitness), int(r.hypothesis)))


def select_best_model(records):
    """Best candidate model for a scene: lowest ICP fitness, ties by model id."""
    ok = [r for r in records if r.converged]
    if not ok:
        return None
    return min(ok, key=lambda r: (float(r.fitness), int(r.model_id)))
