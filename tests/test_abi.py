"""The C-ABI library loads and exports every symbol include/rtr.h declares (no compute calls: no GPU here)."""
import ctypes as C
import os
import re

from conftest import ROOT
from realtime_robot_b200 import _lib
from realtime_robot_b200.params import IcpParams, PoseResult, RansacParams, RegisterParams


def _declared():
    text = open(os.path.join(ROOT, "include", "rtr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rtr_[a-z0-9_]+|ComputeTDFWithCuda)\s*\(", text)))


def test_header_symbols_exported():
    L = _lib.lib()
    names = _declared()
    assert "ComputeTDFWithCuda" in names and "rtr_register" in names and len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_lib.EXPORTS) == names


def test_struct_layouts_match_header():
    assert C.sizeof(PoseResult) == 128
    assert C.sizeof(RansacParams) == 48 and C.sizeof(IcpParams) == 24
    assert C.sizeof(RegisterParams) == 32 + 48 + 24
    p = RegisterParams()
    _lib.lib().rtr_default_register_params(C.byref(p))        # host-only function: safe without a GPU
    from realtime_robot_b200.params import default_register_params
    q = default_register_params()
    assert bytes(p) == bytes(q)
    assert abs(p.harris_radius - 0.05) < 1e-9 and abs(p.harris_threshold - 0.01) < 1e-9   # model_point.h:130-131


def test_reference_ffi_signature_rejects_bad_arguments_without_touching_the_gpu():
    # key_point.h:296,313 can pass num_occ = -1: status code + stderr line, no crash (and no CUDA call needed)
    import numpy as np
    from realtime_robot_b200 import api
    assert api.compute_tdf_with_cuda(np.zeros((0, 3), np.int32), np.zeros(27000, np.float32), 30, -1) != 0
    assert api.compute_tdf_with_cuda(np.zeros((1, 3), np.int32), np.zeros(27000, np.float32), 31, 1) != 0


def test_product_does_not_import_the_oracle():
    """Nothing under realtime_robot_b200/ may link, import, load or execute anything under oracle/ (comments may mention it)."""
    pkg = os.path.join(ROOT, "realtime_robot_b200")
    banned = ("liboracle", "from oracle", "import oracle", "oracle/", "oracle.orc", "orc.", "orc_", "libref_tdf")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for b in banned:
                    assert b not in text, (dirpath, f, b)
