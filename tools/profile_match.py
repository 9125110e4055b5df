#!/usr/bin/env python3
"""Workload for ncu captures of the tensor-core matching kernel: 32768 x 65536 random 33-D descriptors, twice."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from realtime_robot_b200 import api  # noqa: E402

ctx = api.Context(0)
rng = np.random.default_rng(1)
M, N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768, int(sys.argv[2]) if len(sys.argv) > 2 else 65536
fa = (rng.random((M, 33)) * 30).astype(np.float32)
fb = (rng.random((N, 33)) * 30).astype(np.float32)
for _ in range(2):
    ms, st = api.match_raw(ctx, fa, fb, 5)
    print("ms", ms, "redo", st["redo_rows"], "splits", st["splits"], "err", st["observed_err_over_norms"])
ctx.profile_begin()
api.match_raw(ctx, fa, fb, 5)
print({k: round(v[1], 3) for k, v in ctx.profile_end().items()})
