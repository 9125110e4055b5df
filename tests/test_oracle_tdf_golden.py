"""Pins the CPU oracle's TDF to the reference's own kernel (kernel.cu:8-31): tests/golden/tdf_ref.npz holds outputs of
that kernel compiled unmodified and run on a B200 (generator: tools/make_tdf_golden.py)."""
import os

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden", "tdf_ref.npz")


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden file not generated yet")
def test_oracle_tdf_equals_reference_kernel(orc):
    z = np.load(GOLD)
    n = len([k for k in z.files if k.startswith("occ")])
    assert n >= 8
    for k in range(n):
        occ, dim, ref = z[f"occ{k}"], int(z[f"dim{k}"]), z[f"tdf{k}"]
        nv = dim ** 3
        assert np.array_equal(orc.tdf(occ, dim), ref[:nv]), k
        # kernel.cu:13 uses `>`: the reference also writes element dim^3 when it lies inside the 27000 buffer, and
        # nothing beyond it (Appendix B#6)
        if nv < 27000:
            assert ref[nv] != -7.0 and np.all(ref[nv + 1:] == -7.0)
