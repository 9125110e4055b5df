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


# ------------------------------------------------------------------ in-library collective (csrc/comm.cu, NCCL)
def comm_init(ctx, rank: int, world: int, exchange=None):
    """Give `ctx` an NCCL communicator on its own stream (rtr_comm_init).  Rank 0 draws the 128-byte NCCL id
    (rtr_comm_unique_id); `exchange(id_bytes_or_None) -> id_bytes` ships it to the other ranks — default:
    torch.distributed.broadcast_object_list over the already initialised process group (any backend)."""
    from . import _lib
    L = _lib.lib()
    if world <= 1:
        return
    ident = None
    if rank == 0:
        buf = C.create_string_buffer(128)
        _lib.check("rtr_comm_unique_id", L.rtr_comm_unique_id(buf))
        ident = buf.raw
    if exchange is None:
        import torch.distributed as td
        box = [ident]
        td.broadcast_object_list(box, src=0)
        ident = box[0]
    else:
        ident = exchange(ident)
    _lib.check("rtr_comm_init", L.rtr_comm_init(ctx._h, world, rank, ident))


def allgather_results(ctx, records, world: int = None):
    """ONE ncclAllGather of the 128-byte records on the context's stream (rtr_allgather_results): every rank passes the same
    number of records and receives world x that many, in rank order."""
    from . import _lib
    L = _lib.lib()
    n = len(records)
    if world is None:
        w, r = C.c_int(), C.c_int()
        L.rtr_comm_world(ctx._h, C.byref(w), C.byref(r))
        world = w.value
    local = (PoseResult * max(n, 1))(*records)
    out = (PoseResult * max(n * world, 1))()
    _lib.check("rtr_allgather_results", L.rtr_allgather_results(ctx._h, local, n, out))
    return [PoseResult.from_buffer_copy(bytes(out[i])) for i in range(n * world)]


def gather_batches(ctx, on: bool, model_id_base: int = 0):
    """Every rtr_register_many* on `ctx` ends with the in-stream ncclAllGather of its records (rtr_comm_gather_batches)."""
    from . import _lib
    _lib.check("rtr_comm_gather_batches", _lib.lib().rtr_comm_gather_batches(ctx._h, 1 if on else 0, model_id_base))


def gathered_results(ctx, capacity: int = 2048, out=None):
    """The world x n_models records of the last gathered batch, in rank order (rtr_gathered_results), as a ctypes array of
    PoseResult (index it like a list; pass `out` to reuse a buffer)."""
    from . import _lib
    if out is None:
        out = (PoseResult * capacity)()
    n = C.c_int()
    _lib.check("rtr_gathered_results", _lib.lib().rtr_gathered_results(ctx._h, out, len(out), C.byref(n)))
    return out, n.value


def select_best_hypothesis_native(records):
    """rtr_select_best_hypothesis: the library's own arg-min over (fitness, hypothesis id) — what the C++ host calls."""
    from . import _lib
    n = len(records)
    arr = (PoseResult * n)(*records)
    best = PoseResult()
    _lib.check("rtr_select_best_hypothesis", _lib.lib().rtr_select_best_hypothesis(arr, n, C.byref(best)))
    return best


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
