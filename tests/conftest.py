import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def w1():
    from ar_nerf_b200.workload import Workload
    return Workload('W1')


@pytest.fixture(scope="session")
def w3():
    from ar_nerf_b200.workload import Workload
    return Workload('W3', n_poses=20)


def near_clamp(hits_t, near=0.01):
    """rendering.py:31 on a numpy (R,1,2) array; returns the (R,2) view the marchers take."""
    ht = hits_t[:, 0].copy()
    m = (ht[:, 0] >= 0) & (ht[:, 0] < near)
    ht[m, 0] = near
    return ht


def scene_hits(w, rays_o, rays_d):
    import oracle
    s = w.scale
    _, ht, _ = oracle.ray_aabb_intersect(rays_o, rays_d, np.zeros((1, 3), np.float32), np.full((1, 3), s, np.float32), 1)
    return near_clamp(ht)
