"""Oracle neighbour search against brute force and scipy's kd-tree."""
import numpy as np
from scipy.spatial import cKDTree

from conftest import random_cloud


def _sets(counts, offsets, idx):
    return [idx[offsets[i]:offsets[i] + counts[i]] for i in range(len(counts))]


def test_grid_equals_bruteforce(orc, clouds):
    for name, r in (("chair1", 0.05), ("mcloud", 0.1), ("T0_m8111", 0.0365)):
        pts = clouds(name)
        a = orc.radius_neighbors(pts, r, 0)
        b = orc.radius_neighbors(pts, r, 1)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])


def test_radius_semantics(orc):
    # strict d2 < r2 in float, self included, ascending indices
    pts = np.array([[0, 0, 0, 1], [0.05, 0, 0, 1], [0.049999, 0, 0, 1], [0, 0.03, 0.04, 1], [1, 1, 1, 1]], np.float32)
    counts, off, idx = orc.radius_neighbors(pts, 0.05, 0)
    s = _sets(counts, off, idx)
    f = np.float32
    r2 = f(0.05) * f(0.05)
    assert 0 in s[0] and 2 in s[0] and 4 not in s[0]
    assert 1 not in s[0]                   # d2 == r2 exactly: strict compare excludes it
    d2 = f(f(0.03) * f(0.03)) + f(f(0.04) * f(0.04))      # (0, 0.03, 0.04): decided by the float expression
    assert (3 in s[0]) == bool(d2 < r2)
    assert list(s[4]) == [4]


def test_against_ckdtree(orc, clouds):
    pts = clouds("chair2")
    r = 0.05
    counts, off, idx = orc.radius_neighbors(pts, r, 1)
    tree = cKDTree(pts[:, :3].astype(np.float64))
    ref = tree.query_ball_point(pts[:, :3].astype(np.float64), r)
    mism = 0
    for i, s in enumerate(_sets(counts, off, idx)):
        a, b = set(s.tolist()), set(ref[i])
        for j in a ^ b:   # only boundary points may differ (float vs double distance)
            d = np.linalg.norm(pts[i, :3].astype(np.float64) - pts[j, :3].astype(np.float64))
            assert abs(d - r) < 1e-6
            mism += 1
    assert mism < 10


def test_nearest(orc, clouds):
    pts = clouds("desk1")
    rng = np.random.default_rng(3)
    q = np.ones((500, 4), np.float32)
    q[:, :3] = pts[rng.integers(0, len(pts), 500), :3] + rng.normal(0, 0.05, (500, 3)).astype(np.float32)
    q[:50, :3] += 7.0     # far outside the bounding box
    gi, gd = orc.nearest(pts, q, 1)
    bi, bd = orc.nearest(pts, q, 0)
    assert np.array_equal(gi, bi) and np.array_equal(gd, bd)
    di, ii = cKDTree(pts[:, :3].astype(np.float64)).query(q[:, :3].astype(np.float64))
    assert np.mean(ii == gi) > 0.995 and np.allclose(di ** 2, gd, rtol=1e-4, atol=1e-9)


def test_nearest_tie_lowest_index(orc):
    pts = np.array([[1, 0, 0, 1], [-1, 0, 0, 1], [0, 1, 0, 1], [1, 0, 0, 1]], np.float32)
    gi, gd = orc.nearest(pts, np.array([[0, 0, 0, 1]], np.float32), 1)
    assert gi[0] == 0 and gd[0] == 1.0


def test_random_uniform_cloud(orc):
    pts = random_cloud(3000, 5)
    a = orc.radius_neighbors(pts, 0.07, 0)
    b = orc.radius_neighbors(pts, 0.07, 1)
    assert np.array_equal(a[2], b[2])
