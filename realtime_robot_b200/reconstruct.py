"""Scene reconstruction loop (README.md:10,13-17 and img/re.png of the reference): every database
model is registered against every scene segment, the best model per segment is kept and moved into the scene frame.

Pure orchestration over the registration path: rank r registers the models m with m mod W == r (dist.shard_models)
against the replicated segments, one all-gather of 128-byte pose records follows, and every rank takes the same arg-min.
"""
from typing import Dict, List, Sequence

import numpy as np

from . import api, dist
from .params import PoseResult, RegisterParams


def reconstruct(ctx: api.Context, segments: Sequence[np.ndarray], models: Sequence[np.ndarray], params: RegisterParams,
                rank: int = 0, world: int = 1, device=None) -> List[Dict]:
    """segments / models: lists of (n, 4) float32 xyz1 clouds (the same lists on every rank).
    Returns, per segment: {"model": best model id or -1, "pose": 4x4 model -> scene, "fitness", "records": all PoseResult}."""
    seg_d = [api.Cloud(ctx, s) for s in segments]
    mine = dist.shard_models(len(models), rank, world)
    mod_d = {m: api.Cloud(ctx, models[m]) for m in mine}
    per_rank = (len(models) + world - 1) // world
    out = []
    for si, sd in enumerate(seg_d):
        recs = []
        for m in mine:
            mod_d[m].reset(); sd.reset()                 # no stage is carried over from another pairing
            r = api.register(mod_d[m], sd, params)
            r.model_id = m
            recs.append(r)
        allr = dist.all_gather_records(recs, per_rank, device=device) if world > 1 else recs
        best = dist.select_best_model(allr)
        out.append({"segment": si, "model": -1 if best is None else int(best.model_id),
                    "pose": np.eye(4, dtype=np.float32) if best is None else best.matrix(),
                    "fitness": float("inf") if best is None else float(best.fitness), "records": allr})
    for c in seg_d:
        c.free()
    for c in mod_d.values():
        c.free()
    return out


def compose_scene(segments_result: List[Dict], models: Sequence[np.ndarray]) -> np.ndarray:
    """The reconstructed scene: every winning model transformed into the scene frame (what img/re.png shows)."""
    parts = []
    for r in segments_result:
        if r["model"] < 0:
            continue
        m = models[r["model"]]
        T = r["pose"].astype(np.float64)
        p = np.ones((len(m), 4), dtype=np.float32)
        p[:, :3] = (m[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
        parts.append(p)
    return np.concatenate(parts) if parts else np.zeros((0, 4), dtype=np.float32)
