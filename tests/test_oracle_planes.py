"""Oracle restatement of ModelPoint::getArea (oracle/planes.cpp): RANSAC plane peel + convex-hull area.
PCL internals from SURVEY App. A.7 ([PCL-mem]): parity unpinned; the hull areas are cross-checked against real Qhull (scipy)."""
import numpy as np
from scipy.spatial import ConvexHull

from realtime_robot_b200 import synth


def test_mt19937_draws_match_boost_convention():
    # boost::mt19937(12345u) == std::mt19937(12345) == numpy's legacy init_genrand seeding; rnd() = raw >> 1
    raw = np.random.RandomState(12345)._bit_generator.random_raw(3)
    assert [int(v) for v in raw] == [3992670690, 3823185381, 1358822685]


def test_hull_area_equals_qhull(orc):
    rng = np.random.default_rng(0)
    # horizontal rectangle 0.5 x 0.4 with mm noise: 2-D, projected on xy
    p = np.ones((2000, 4), np.float32)
    p[:, 0], p[:, 1], p[:, 2] = rng.random(2000) * 0.5, rng.random(2000) * 0.4, 0.3 + rng.normal(0, 0.001, 2000)
    a, dim = orc.hull_area(p)
    assert dim == 2 and abs(a - ConvexHull(p[:, :2].astype(np.float64)).volume) < 1e-12 and abs(a - 0.2) < 0.005
    # vertical plane with normal along x: projected on yz
    q = p.copy(); q[:, 0], q[:, 2] = p[:, 2], p[:, 0]
    a, dim = orc.hull_area(q)
    assert dim == 2 and abs(a - ConvexHull(q[:, 1:3].astype(np.float64)).volume) < 1e-12
    # vertical plane at 45 degrees between x and y: no axis within 10 degrees -> xy projection -> a sliver (the PCL quirk)
    r = p.copy(); r[:, 0] = (p[:, 0] + p[:, 2]) * np.float32(0.7071); r[:, 1] = (p[:, 0] - p[:, 2]) * np.float32(0.7071); r[:, 2] = p[:, 1]
    a, dim = orc.hull_area(r)
    assert dim == 2 and abs(a - ConvexHull(r[:, :2].astype(np.float64)).volume) < 1e-12 and a < 0.01
    # a thick blob: 3-D mode -> twice the best-fit-plane hull area; Qhull's facet area is close to that for a slab
    s = np.ones((3000, 4), np.float32); s[:, :3] = (rng.random((3000, 3)) * np.array([0.1, 0.1, 0.01])).astype(np.float32)
    a, dim = orc.hull_area(s)
    assert dim == 3 and abs(a - ConvexHull(s[:, :3].astype(np.float64)).area) / a < 0.25
    assert orc.hull_area(p[:2])[0] == 0.0


def test_plane_segment_on_a_box(orc):
    # floor 1 x 1 plus a 0.4 m wall: the first plane RANSAC finds is the floor (most inliers)
    rects = [(np.zeros(3), np.array([1.0, 0, 0]), np.array([0, 1.0, 0])), (np.zeros(3), np.array([1.0, 0, 0]), np.array([0, 0, 0.4]))]
    pts = synth.sample_rects(rects, 7000, 3, noise=0.0005)
    coeff, idx, its = orc.plane_segment(pts)
    assert abs(abs(coeff[2]) - 1) < 1e-3 and abs(coeff[3]) < 2e-3 and its <= 151
    on_floor = np.abs(pts[:, 2]) < 0.004
    assert np.mean(on_floor[idx]) > 0.99 and len(idx) > 0.95 * on_floor.sum()
    assert orc.lib().orc_plane_class(coeff.ctypes.data_as(__import__("ctypes").c_void_p)) == 1


def test_plane_areas_repo_clouds(orc, clouds):
    s = orc.plane_areas(clouds("chair1"))
    assert len(s) == 14 and sum(x.inliers for x in s) >= 0.85 * 1818
    seat = s[0]
    assert seat.kept == 1 and seat.is_vertical == 0 and abs(seat.coefficients[2]) > 0.99 and 0.16 <= seat.area < 0.2
    assert all(x.dimension in (2, 3) and x.iterations <= 151 for x in s)
    kept = [x for x in s if x.kept]
    assert all(x.area >= 0.16 for x in kept) and len(kept) == 7
    # Appendix B#1: after the as-committed x0.01 scale every plane is far below the 0.16 m^2 gate -> no surfaces
    tiny = clouds("chair1").copy(); tiny[:, :3] *= np.float32(0.01)
    assert not any(x.kept for x in orc.plane_areas(tiny))
