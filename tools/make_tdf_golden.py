#!/usr/bin/env python3
"""Golden vectors for the TDF stage: outputs of the REFERENCE's own kernel.cu (compiled unmodified into
oracle/_ref/libref_tdf.so by oracle/Makefile) on seeded inputs.  Must run on a GPU box (the reference code is CUDA):
    gpurun -- python tools/make_tdf_golden.py gpurun_out/tdf_ref.npz
The result is committed as tests/golden/tdf_ref.npz; tests/test_oracle_tdf_golden.py pins the CPU oracle to it."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def cases():
    rng = np.random.default_rng(20170427)
    out = []
    for n_occ, dim in [(0, 30), (1, 30), (2, 30), (35, 30), (191, 30), (777, 30), (60, 12), (5, 1), (300, 29)]:
        hi = dim + 2 if n_occ % 2 else dim           # some lists reach outside [0, dim) like key_point.h:300-307 can
        out.append((rng.integers(-1 if n_occ % 2 else 0, hi, size=(n_occ, 3)).astype(np.int32), dim))
    return out


def main(dst):
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_tdf.so"))
    data = {}
    for k, (occ, dim) in enumerate(cases()):
        buf = np.full(27000, -7.0, np.float32)        # sentinel: shows which elements the reference writes
        occ_c = np.ascontiguousarray(occ if len(occ) else np.zeros((1, 3), np.int32))
        rc = ref.ComputeTDFWithCuda(occ_c.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p), dim, len(occ))
        assert rc == 0, rc
        data[f"occ{k}"] = occ
        data[f"dim{k}"] = np.int32(dim)
        data[f"tdf{k}"] = buf
    np.savez_compressed(dst, **data)
    print("wrote", dst, len(cases()), "cases")


if __name__ == "__main__":
    main(sys.argv[1])
