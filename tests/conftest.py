import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def cloud_path(name):
    return os.path.join(ROOT, "data", "clouds", name + ".pcd")


@pytest.fixture(scope="session")
def clouds():
    from realtime_robot_b200.pcd import read_pcd_xyz, to_xyz1
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = to_xyz1(read_pcd_xyz(cloud_path(name)))
        return cache[name]
    return get


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as o
    o.build()
    return o


@pytest.fixture(scope="session")
def gpu_ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from realtime_robot_b200 import api
    return api.Context(0)


def random_cloud(n, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    out = np.ones((n, 4), dtype=np.float32)
    out[:, :3] = (rng.random((n, 3)) * scale).astype(np.float32)
    return out
