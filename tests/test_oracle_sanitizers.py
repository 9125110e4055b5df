"""The CPU oracle built with -fsanitize=address,undefined (oracle/Makefile target oracle_asan) and driven over every stage of
the path on the chair1 / mcloud pair, plus empty and 2-point clouds: SURVEY section 5 rows 1-2 (memory-error and UB detection
for the checker itself — a checker that reads out of bounds pins nothing)."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_is_clean_under_asan_and_ubsan(clouds, tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle_asan"])
    paths = []
    for name in ("chair1", "mcloud"):
        p = tmp_path / (name + ".f32")
        np.ascontiguousarray(clouds(name), dtype=np.float32).tofile(p)
        paths.append(str(p))
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1", OMP_NUM_THREADS="2")
    r = subprocess.run([os.path.join(ROOT, "oracle", "oracle_asan")] + paths, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "asan ok" in r.stdout
    assert "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]
