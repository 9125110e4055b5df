"""Oracle normals / Harris / FPFH / feature matching against closed forms and numpy."""
import numpy as np

from realtime_robot_b200 import synth


def _plane(n=40, step=0.01, z=0.3):
    g = np.arange(n) * step
    x, y = np.meshgrid(g, g)
    p = np.ones((n * n, 4), np.float32)
    p[:, 0], p[:, 1], p[:, 2] = x.ravel(), y.ravel(), z
    return p


def test_normals_plane(orc):
    p = _plane()
    n4 = orc.normals(p, 0.035)
    assert np.allclose(np.abs(n4[:, 2]), 1.0, atol=1e-6) and np.allclose(n4[:, :2], 0, atol=1e-6)
    assert np.all(n4[:, 2] < 0)            # flipped towards the viewpoint (0,0,0): n . (0 - p) >= 0 with z > 0
    assert np.allclose(n4[:, 3], 0, atol=1e-9)


def test_normals_few_neighbours_nan(orc):
    p = np.array([[0, 0, 0, 1], [0.01, 0, 0, 1], [5, 5, 5, 1]], np.float32)
    n4 = orc.normals(p, 0.05)
    assert np.all(np.isnan(n4))


def test_normals_match_numpy_eigh(orc, clouds):
    pts = clouds("chair1")
    n4 = orc.normals(pts, 0.05)
    counts, off, idx = orc.radius_neighbors(pts, 0.05, 1)
    P = pts[:, :3].astype(np.float64)
    worst = 0.0
    for i in range(0, len(pts), 7):
        nb = P[idx[off[i]:off[i] + counts[i]]]
        w, v = np.linalg.eigh(np.cov(nb.T, bias=True))
        gap = (w[1] - w[0]) / max(w[2], 1e-30)
        if gap < 1e-3:
            continue                      # direction ill-defined
        worst = max(worst, 1.0 - abs(float(v[:, 0] @ n4[i, :3].astype(np.float64))))
        assert abs(n4[i, 3] - w[0] / w.sum()) < 1e-6
    assert worst < 3e-7                   # float32 storage of the unit vector


def test_normals_pcl_float_mode_is_close(orc, clouds):
    pts = clouds("chair1")
    a, b = orc.normals(pts, 0.05, 0), orc.normals(pts, 0.05, 1)
    dots = np.abs(np.sum(a[:, :3] * b[:, :3], axis=1))
    assert np.median(1 - dots) < 1e-7 and np.mean(1 - dots < 1e-3) > 0.97


def test_harris_response_cube_corner(orc):
    # three orthogonal unit-normal families in equal numbers: C = I/3 -> det = 1/27, trace = 1 -> response = 1/27
    pts = np.ones((30, 4), np.float32)
    pts[:, :3] = np.random.default_rng(0).random((30, 3)).astype(np.float32) * 0.01
    n4 = np.zeros((30, 4), np.float32)
    for k in range(30):
        n4[k, k % 3] = 1.0
    resp, ki, kx = orc.harris3d(pts, n4, 0.05, 0.01, 1, 0)
    assert np.allclose(resp, 1.0 / 27.0, atol=1e-7)
    assert len(ki) == 30                   # equal responses: nobody is strictly larger, all survive


def test_harris_plane_has_no_corners(orc):
    p = _plane()
    resp, ki, _ = orc.harris3d(p, orc.normals(p, 0.035), 0.035, 0.01)
    assert len(ki) == 0 and np.abs(resp).max() < 1e-9


def test_harris_repo_clouds(orc, clouds):
    # SURVEY 8(a1) row 1 [emu]: 7 / 14 / 11 corners on chair1 / mcloud / T0_m8111 with r = 0.05, thr = 0.01
    for name, k in (("chair1", 7), ("mcloud", 14), ("T0_m8111", 11)):
        pts = clouds(name)
        resp, ki, kx = orc.harris3d(pts, orc.normals(pts, 0.05), 0.05, 0.01)
        assert len(ki) == k, (name, len(ki))
        assert np.all(np.diff(ki) > 0) and np.all(resp[ki] >= 0.01)
        assert resp.max() <= 1.0 / 27.0 + 1e-6


def test_harris_scaled_model_has_no_keypoints(orc, clouds):
    # Appendix B#1: the as-committed in-place x0.01 scale (model_point.h:105-111) leaves every point in everyone's
    # neighbourhood -> response ~ 0 -> zero corners.
    pts = clouds("chair1").copy()
    pts[:, :3] *= np.float32(0.01)
    resp, ki, _ = orc.harris3d(pts, orc.normals(pts, 0.05), 0.05, 0.01)
    assert len(ki) == 0


def _fpfh_numpy(pts, n4, r):
    """Independent, slow restatement of App. A.4 for small clouds."""
    P, N = pts[:, :3].astype(np.float64), n4[:, :3].astype(np.float64)
    n = len(P)
    d2f = lambda a, b: np.float32(np.float32(np.float32(a[0] - b[0]) ** 2 + np.float32(a[1] - b[1]) ** 2) + np.float32(a[2] - b[2]) ** 2)
    nbrs = [[j for j in range(n) if d2f(pts[i], pts[j]) < np.float32(r) * np.float32(r)] for i in range(n)]
    spfh = np.zeros((n, 33), np.float32)
    for i in range(n):
        if len(nbrs[i]) < 2 or not np.all(np.isfinite(N[i])):
            continue
        cnt = np.zeros(33)
        for j in nbrs[i]:
            if j == i or not np.all(np.isfinite(N[j])):
                continue
            d = P[j] - P[i]
            f4 = np.linalg.norm(d)
            if f4 == 0:
                continue
            n1, n2 = N[i], N[j]
            a1, a2 = n1 @ d / f4, n2 @ d / f4
            if abs(a1) < abs(a2):
                n1, n2, d, f3 = n2, n1, -d, -a2
            else:
                f3 = a1
            v = np.cross(d, n1)
            if np.linalg.norm(v) == 0:
                continue
            v /= np.linalg.norm(v)
            w = np.cross(n1, v)
            f2, f1 = v @ n2, np.arctan2(w @ n2, n1 @ n2)
            b = [int(np.floor(11 * (f1 + np.pi) / (2 * np.pi))), int(np.floor(11 * (f2 + 1) / 2)), int(np.floor(11 * (f3 + 1) / 2))]
            for t in range(3):
                cnt[t * 11 + min(max(b[t], 0), 10)] += 1
        spfh[i] = (cnt * (100.0 / (len(nbrs[i]) - 1))).astype(np.float32)
    out = np.zeros((n, 33), np.float32)
    for i in range(n):
        h = np.zeros(33)
        for j in nbrs[i]:
            dd = d2f(pts[i], pts[j])
            if dd == 0:
                continue
            h += spfh[j].astype(np.float64) / float(dd)
        for t in range(3):
            s = h[t * 11:(t + 1) * 11].sum()
            out[i, t * 11:(t + 1) * 11] = (h[t * 11:(t + 1) * 11] * (100.0 / s if s != 0 else 0.0)).astype(np.float32)
    return out


def test_fpfh_against_numpy(orc):
    pts = synth.sample_rects(synth.chair_rects(), 160, 11)
    n4 = orc.normals(pts, 0.12)
    a = orc.fpfh(pts, n4, 0.2)
    b = _fpfh_numpy(pts, n4, 0.2)
    assert np.abs(a - b).max() < 2e-3      # numpy sums in a different order; a bin flip would show as >> 1


def test_fpfh_properties(orc, clouds):
    pts = clouds("chair1")
    n4 = orc.normals(pts, 0.05)
    f = orc.fpfh(pts, n4, 0.10)
    assert f.shape == (1818, 33) and np.all(f >= 0)
    for t in range(3):
        s = f[:, t * 11:(t + 1) * 11].sum(1)
        assert np.allclose(s[s > 0], 100.0, atol=1e-3)
    # rigid invariance (up to float noise in the transformed coordinates / viewpoint flip): translate only, away from 0
    # so that no normal flips: FPFH uses angles between normals and offsets, both unchanged.
    moved = pts.copy(); moved[:, :3] += np.float32(0.25)
    n4m = orc.normals(moved, 0.05)
    same_sign = np.sum(n4[:, :3] * n4m[:, :3], axis=1) > 0
    fm = orc.fpfh(moved, n4m, 0.10)
    assert np.median(np.abs(fm - f).max(1)[same_sign]) < 6.0


def test_match_features_bruteforce(orc):
    rng = np.random.default_rng(4)
    fa, fb = rng.random((70, 33)).astype(np.float32) * 100, rng.random((300, 33)).astype(np.float32) * 100
    fb[17] = fb[5]                          # exact tie: lowest index first
    idx, dist = orc.match_features(fa, fb, 5)
    d = ((fa[:, None, :].astype(np.float64) - fb[None].astype(np.float64)) ** 2).sum(-1).astype(np.float32)
    order = np.lexsort((np.broadcast_to(np.arange(300), d.shape), d), axis=1)[:, :5]
    assert np.array_equal(idx, order)
    assert np.allclose(dist, np.take_along_axis(d, order, 1), rtol=1e-6)
    assert np.all(np.diff(dist, axis=1) >= 0)
