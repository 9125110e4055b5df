"""Oracle pose stages: Jacobi / Horn solves, prerejective RANSAC (App. A.5), ICP (App. A.6), TDF (kernel.cu)."""
import numpy as np

from realtime_robot_b200 import synth
from realtime_robot_b200.params import default_register_params


def kabsch(S, T):
    cs, ct = S.mean(0), T.mean(0)
    H = (S - cs).T @ (T - ct)
    U, _, Vt = np.linalg.svd(H)
    D = np.diag([1, 1, np.sign(np.linalg.det(Vt.T @ U.T))])
    R = Vt.T @ D @ U.T
    M = np.eye(4); M[:3, :3] = R; M[:3, 3] = ct - R @ cs
    return M


def test_jacobi_matches_eigh(orc):
    rng = np.random.default_rng(0)
    for n in (3, 4):
        for _ in range(50):
            a = rng.standard_normal((n, n)); a = a + a.T
            ev, vec = orc.jacobi(a)
            w, v = np.linalg.eigh(a)
            assert np.allclose(np.sort(ev), w, atol=1e-12)
            assert np.allclose(vec @ np.diag(ev) @ vec.T, a, atol=1e-12)
    ev, vec = orc.jacobi(np.diag([3.0, 1.0, 2.0]))
    assert np.array_equal(ev, [3.0, 1.0, 2.0]) and np.array_equal(vec, np.eye(3))


def _horn_matrix(S):
    N = np.zeros((4, 4))
    N[0, 0] = S[0, 0] + S[1, 1] + S[2, 2]
    N[0, 1] = S[1, 2] - S[2, 1]; N[0, 2] = S[2, 0] - S[0, 2]; N[0, 3] = S[0, 1] - S[1, 0]
    N[1, 1] = S[0, 0] - S[1, 1] - S[2, 2]; N[1, 2] = S[0, 1] + S[1, 0]; N[1, 3] = S[2, 0] + S[0, 2]
    N[2, 2] = -S[0, 0] + S[1, 1] - S[2, 2]; N[2, 3] = S[1, 2] + S[2, 1]; N[3, 3] = -S[0, 0] - S[1, 1] + S[2, 2]
    return N + np.triu(N, 1).T


def test_horn_quartic_route_matches_eigh(orc):
    """horn_pose's fast path (top root of the characteristic quartic + adjugate column) against numpy's eigh on random,
    3-point, near-identity (ICP) and planar cross-covariances; collinear inputs (repeated top eigenvalue) must decline."""
    rng = np.random.default_rng(0)
    declined = 0
    for t in range(4000):
        kind = t % 5
        if kind == 0:
            S = rng.normal(size=(3, 3)) * 10 ** rng.uniform(-4, 6)
        elif kind == 1:
            s = rng.normal(size=(3, 3)); R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
            tt = s @ R.T + rng.normal(size=(3, 3)) * 0.01
            S = (s - s.mean(0)).T @ (tt - tt.mean(0))
        elif kind == 2:
            s = rng.normal(size=(500, 3)); tt = s + rng.normal(size=(500, 3)) * 1e-3
            S = (s - s.mean(0)).T @ (tt - tt.mean(0))
        elif kind == 3:
            s = np.outer(rng.normal(size=50), rng.normal(size=3)) + np.outer(rng.normal(size=50), rng.normal(size=3))
            S = s.T @ s
        else:
            s = np.outer(rng.normal(size=50), rng.normal(size=3))
            S = s.T @ s
        N = _horn_matrix(S)
        ok, q = orc.horn_top_eigvec(N)
        if kind == 4:
            assert not ok
            declined += 1
            continue
        assert ok
        v = np.linalg.eigh(N)[1][:, -1]
        q = q / np.linalg.norm(q)
        assert min(np.linalg.norm(q - v), np.linalg.norm(q + v)) < 1e-7
    assert declined == 800
    assert orc.horn_top_eigvec(np.zeros((4, 4)))[0] is False


def test_horn_matches_kabsch(orc, clouds):
    src = clouds("chair1")[:500]
    gt = synth.rigid(20, -35, 110, (0.3, -0.2, 0.5), about=(0.1, 0.2, 0.3))
    tgt = synth.apply(gt, src)
    tgt[:, :3] += np.random.default_rng(1).normal(0, 0.002, (500, 3)).astype(np.float32)
    M = orc.pose_from_pairs(src, tgt)
    K = kabsch(src[:, :3].astype(np.float64), tgt[:, :3].astype(np.float64))
    assert np.abs(M - K).max() < 2e-6
    assert abs(np.linalg.det(M[:3, :3].astype(np.float64)) - 1) < 1e-6
    # three points (always coplanar -> rank-2 cross covariance): still the proper rotation
    M3 = orc.pose_from_pairs(src[:3], synth.apply(gt, src[:3]))
    assert np.abs(M3 - gt).max() < 1e-4


def test_transform_matches_definition(orc):
    rng = np.random.default_rng(2)
    p = np.ones((100, 4), np.float32); p[:, :3] = rng.standard_normal((100, 3)).astype(np.float32)
    M = synth.rigid(10, 20, 30, (1, 2, 3)).astype(np.float32)
    out = orc.transform(p, M)
    f = np.float32
    exp = np.stack([f(f(f(M[r, 0] * p[:, 0]) + f(M[r, 1] * p[:, 1])) + f(M[r, 2] * p[:, 2])) + M[r, 3] for r in range(3)], 1)
    assert np.array_equal(out[:, :3], exp.astype(np.float32))


def test_tdf_matches_numpy(orc):
    rng = np.random.default_rng(3)
    for n_occ, dim in ((0, 30), (1, 30), (191, 30), (40, 12)):
        occ = rng.integers(0, dim + 1, (n_occ, 3)).astype(np.int32)
        out = orc.tdf(occ, dim)
        z, y, x = np.meshgrid(np.arange(dim), np.arange(dim), np.arange(dim), indexing="ij")
        g = np.stack([x, y, z], -1).reshape(-1, 1, 3)
        exp = np.minimum(900, ((g - occ[None]) ** 2).sum(-1).min(1)) if n_occ else np.full(dim ** 3, 900)
        assert np.array_equal(out, exp.astype(np.float32))      # layout idx = z*dim^2 + y*dim + x (kernel.cu:16-18)


def test_ransac_known_pose(orc, clouds):
    # apply a known rigid motion to chair1 and register it back
    model = clouds("chair1")
    gt = synth.rigid(8, -5, 40, (0.4, -0.3, 0.2), about=(0.2, 0.2, 0.4))
    scene = synth.apply(gt, model)
    p = default_register_params()
    p.ransac.max_iterations = 4000
    nm, ns = orc.normals(model, 0.05), orc.normals(scene, 0.05)
    knn, _ = orc.match_features(orc.fpfh(model, nm, 0.1), orc.fpfh(scene, ns, 0.1), 5)
    res = orc.ransac(model, scene, knn, p.ransac)
    assert res.converged == 1 and res.hypothesis >= 0 and 0 < res.evaluated < 4000
    assert res.inliers > 0.9 * len(model)
    assert np.abs(res.matrix() - gt).max() < 0.02
    # hypothesis sharding: the union of two half ranges gives the same winner (8e)
    a, b = default_register_params().ransac, default_register_params().ransac
    a.max_iterations = b.max_iterations = 4000
    a.hypothesis_begin, a.hypothesis_end = 0, 2000
    b.hypothesis_begin, b.hypothesis_end = 2000, 4000
    ra, rb = orc.ransac(model, scene, knn, a), orc.ransac(model, scene, knn, b)
    assert ra.evaluated + rb.evaluated == res.evaluated
    best = min((r for r in (ra, rb) if r.hypothesis >= 0), key=lambda r: (r.fitness, r.hypothesis))
    assert best.hypothesis == res.hypothesis and np.array_equal(best.matrix(), res.matrix())
    # one hypothesis, inspected
    ok, s6, pose = orc.hypothesis(model, scene, knn, p.ransac, res.hypothesis)
    assert ok == 1 and len(set(s6[:3].tolist())) == 3 and np.array_equal(pose, res.matrix())


def test_ransac_no_solution(orc, clouds):
    model = clouds("chair1")
    p = default_register_params()
    p.ransac.max_iterations = 200
    p.ransac.inlier_fraction = 1.1          # unattainable
    scene = clouds("mcloud")
    knn, _ = orc.match_features(orc.fpfh(model, orc.normals(model, 0.05), 0.1), orc.fpfh(scene, orc.normals(scene, 0.05), 0.1), 5)
    res = orc.ransac(model, scene, knn, p.ransac)
    assert res.hypothesis == -1 and res.converged == 0 and np.array_equal(res.matrix(), np.eye(4, dtype=np.float32))


def test_icp_known_translation_pair(orc, clouds):
    # the reference ships 70761_c = 70761 + (1, 0.5, 0); start ICP near it and it must land on the exact translation
    src, tgt = clouds("70761"), clouds("70761_c")
    p = default_register_params()
    p.icp.max_iterations = 50
    init = np.eye(4); init[:3, 3] = (0.99, 0.505, 0.005)
    res = orc.icp(src, tgt, p.icp, init)
    assert res.converged in (2, 3) and res.iterations < 10
    exp = np.eye(4); exp[:3, 3] = (1.0, 0.5, 0.0)
    assert np.abs(res.matrix() - exp).max() < 1e-5 and res.fitness < 1e-10


def test_icp_defaults_and_states(orc, clouds):
    src = clouds("chair1")
    tgt = synth.apply(synth.rigid(1, 2, 4, (0.01, 0.02, -0.01), about=(0.2, 0.2, 0.4)), src)
    p = default_register_params()
    res = orc.icp(src, tgt, p.icp)           # PCL defaults: 10 iterations
    assert res.iterations == 10 and res.converged == 1
    p.icp.max_iterations = 3; p.icp.force_iterations = 1
    assert orc.icp(src, tgt, p.icp).iterations == 3
    # fewer than three correspondences -> not converged, pose untouched
    p = default_register_params(); p.icp.max_correspondence_distance = 1e-6
    far = tgt.copy(); far[:, :3] += 3
    r = orc.icp(src, far, p.icp)
    assert r.converged == 0 and r.iterations == 0 and np.array_equal(r.matrix(), np.eye(4, dtype=np.float32))


def test_point_to_plane_icp_recovers_pose_in_fewer_iterations(orc, clouds):
    """estimator 1 (TransformationEstimationPointToPlaneLLS restated: 6x6 normal equations, Gaussian elimination,
    constructTransformationMatrix): on an exact rigid copy it converges to the ground truth, faster than the SVD estimator."""
    from realtime_robot_b200.params import default_register_params
    m = clouds("chair1")
    gt = synth.rigid(3, -2, 4, (0.02, -0.015, 0.01), about=(0.2, 0.2, 0.4))
    s = synth.apply(gt, m)
    n4 = orc.normals(s, 0.05)
    p = default_register_params()
    p.icp.max_iterations = 30
    res = {}
    for est in (0, 1):
        p.icp.estimator = est
        res[est] = orc.icp(m, s, p.icp, None, n4)
        assert res[est].converged and np.abs(res[est].matrix() - gt).max() < 1e-5
    assert res[1].iterations < res[0].iterations


def test_point_to_plane_step_matches_numpy_normal_equations(orc, clouds):
    """One estimator-1 iteration == numpy: exact 1-NN pairs (cKDTree), rows [s x n ; n], right-hand sides n.(t - s),
    numpy.linalg.solve of the 6x6 normal equations, R = Rz(gamma) Ry(beta) Rx(alpha) (constructTransformationMatrix)."""
    from scipy.spatial import cKDTree
    tgt = clouds("chair1")
    gt = synth.rigid(2, -1, 3, (0.01, -0.008, 0.006), about=(0.2, 0.2, 0.4))
    src = synth.apply(np.linalg.inv(gt), tgt)[::3].copy()
    n4 = orc.normals(tgt, 0.05)
    p = default_register_params()
    p.icp.max_iterations = 1; p.icp.estimator = 1
    r = orc.icp(src, tgt, p.icp, None, n4)
    assert r.iterations == 1
    _, nn = cKDTree(tgt[:, :3].astype(np.float64)).query(src[:, :3].astype(np.float64))
    s, t, n = src[:, :3].astype(np.float64), tgt[nn, :3].astype(np.float64), n4[nn, :3].astype(np.float64)
    ok = np.isfinite(n).all(1)
    s, t, n = s[ok], t[ok], n[ok]
    A = np.hstack([np.cross(s, n), n])                     # linearised (R s + t - d) . n: omega . (s x n) + t . n = (d - s) . n
    b = (n * (t - s)).sum(1)
    x = np.linalg.solve(A.T @ A, A.T @ b)
    ca, sa, cb, sb, cg, sg = np.cos(x[0]), np.sin(x[0]), np.cos(x[1]), np.sin(x[1]), np.cos(x[2]), np.sin(x[2])
    Rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]]); Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]]); Rz = np.array([[cg, -sg, 0], [sg, cg, 0], [0, 0, 1]])
    M = np.eye(4); M[:3, :3] = Rz @ Ry @ Rx; M[:3, 3] = x[3:]
    assert np.abs(r.matrix() - M).max() < 1e-6
    R = r.matrix()[:3, :3].astype(np.float64)
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-6 and abs(np.linalg.det(R) - 1.0) < 1e-6
