"""Oracle restatement of the reference's own descriptor path (oracle/native.cpp): occupancy grid, TDF voxel list,
36-step yaw sweep, screens, exhaustive consensus.  Octree semantics are [PCL-mem]: parity unpinned."""
import numpy as np

from realtime_robot_b200 import synth
from realtime_robot_b200.params import default_native_params


def _kps(orc, pts):
    return orc.harris3d(pts, orc.normals(pts, 0.05), 0.05, 0.01)[2]


def test_occupancy_matches_survey_ranges(orc, clouds):
    # SURVEY 8(a1) row 4 [emu]: 37-192 points and 35-191 occupied voxels per chair1 keypoint box
    pts = clouds("chair1")
    kp = _kps(orc, pts)
    number, count, tdf, vox, occ = orc.native_keypoint_descriptors(pts, kp, default_native_params())
    live = count > 0
    assert count[live].min() == 37 and count[live].max() == 192
    assert number[live].min() == 35 and number[live].max() == 191
    for k in np.flatnonzero(live):
        d = np.abs(occ[k][:, :3] - kp[k, :3]).max()
        assert d <= 0.1 + 1e-6
        # brute force: every point of the cloud inside the inclusive float box is in the occupancy cloud
        lo, hi = kp[k, :3] - np.float32(0.1), kp[k, :3] + np.float32(0.1)
        inside = np.all((pts[:, :3] >= lo) & (pts[:, :3] <= hi), axis=1)
        assert inside.sum() == count[k]
        # Number = distinct 1 cm voxels in the frame origin = min - 0.06 (App. A.8: +-0.1 box enlarged to 32 voxels)
        origin = lo.astype(np.float64) - (32 * np.float64(np.float32(0.01)) - (hi.astype(np.float64) - lo.astype(np.float64))) / 2
        keys = np.floor((occ[k][:, :3].astype(np.float64) - origin) / np.float64(np.float32(0.01))).astype(int)
        assert len({tuple(r) for r in keys}) == number[k]
        assert vox[k] == number[k]          # the +-0.15 frame has the same origin (c - 0.16): same voxel set, nothing skipped
        assert np.all(tdf[k] >= 0) and tdf[k].max() <= 900 and (tdf[k] == 0).sum() >= 1


def test_tdf_voxel_quirks(orc, clouds):
    pts = clouds("chair1")
    kp = _kps(orc, pts)[:3]
    p = default_native_params()
    a = orc.native_keypoint_descriptors(pts, kp, p)
    p.quirk_skip_first_voxel = 1
    b = orc.native_keypoint_descriptors(pts, kp, p)
    assert np.array_equal(b[3], a[3] - 1)                  # key_point.h:298 starts at i = 1: one voxel fewer (B#5)
    assert np.all(b[2] >= a[2])                            # fewer occupied voxels can only raise the distance field


def test_pair_score_recovers_yaw(orc, clouds):
    # rotate chair1 by 40 degrees about z and shift it: corresponding keypoints must score best at step 32 (320 = -40 deg)
    model = clouds("chair1")
    scan = synth.apply(synth.rigid(0, 0, 40, (0.3, -0.2, 0.0), about=(0.2, 0.2, 0.0)), model)
    p = default_native_params()
    mk, sk = _kps(orc, model), _kps(orc, scan)
    assert len(mk) == len(sk) == 7
    dm = orc.native_keypoint_descriptors(model, mk, p)
    ds = orc.native_keypoint_descriptors(scan, sk, p, with_tdf=False)
    hits = 0
    for k in range(6):                                     # the 7th refined corner drifted off the cloud (empty box)
        sc = [orc.native_pair_score(mk[k], dm[2][k], ds[4][s], sk[s], p) for s in range(7)]
        best = int(np.argmin([v[0] for v in sc]))
        if best == k and sc[k][1] == 32:
            hits += 1
            T = sc[k][2]
            assert np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3) * (mk[k, 2] / sk[k, 2]) ** 2, atol=1e-4)    # rotation x z-ratio scale
    assert hits >= 5
    s, step, T = orc.native_pair_score(mk[6], dm[2][6], ds[4][0], sk[0], p)
    assert s == 810000.0                                   # empty TDF (all 900): every voxel scores 900^2


def test_screens_and_consensus(orc, clouds):
    model = clouds("chair1")
    scan = synth.apply(synth.rigid(0, 0, 40, (0.3, -0.2, 0.0), about=(0.2, 0.2, 0.0)), model)
    mk, sk = _kps(orc, model), _kps(orc, scan)
    p = default_native_params()
    r = orc.native_register(model, mk, scan, sk, p)
    assert r.evaluated == 0 and r.hypothesis == -1 and np.array_equal(r.matrix(), np.eye(4, dtype=np.float32))   # gate < 3 passes nothing here
    p.pair_gate = 30.0                                     # admit the true pairs (their scores are 7..22)
    r = orc.native_register(model, mk, scan, sk, p)
    assert r.converged == 1 and r.inliers >= 4 and r.evaluated >= r.inliers
    T = r.matrix().astype(np.float64)
    yaw = np.degrees(np.arctan2(T[1, 0], T[0, 0]))
    assert abs(yaw - (-40.0)) < 1e-3 or abs(yaw - 320.0) < 1e-3
    # the as-committed screens: integer division and float(2/3) == 0 (function.h:161,174)
    L = orc.lib()
    import ctypes as C
    a = np.array([0, 0, 0.9, 1], np.float32); b = np.array([0, 0, 0.5, 1], np.float32)
    ar = (C.c_double * 3)(0.16, 0.16, 0.16)
    fp = lambda x: x.ctypes.data_as(C.c_void_p)
    q = default_native_params()
    assert L.orc_native_screens(fp(a), fp(b), ar, ar, 100, 60, C.byref(q)) == 0      # height ratio 1.8 > 1.5
    q.quirk_integer_screens = 1
    assert L.orc_native_screens(fp(a), fp(b), ar, ar, 100, 60, C.byref(q)) == 1      # height always passes; 100/60 -> 1
    assert L.orc_native_screens(fp(a), fp(b), ar, ar, 59, 60, C.byref(q)) == 0       # 59/60 -> 0 < 0.5
    q.quirk_integer_screens = 0
    assert L.orc_native_screens(fp(b), fp(b), ar, ar, 59, 60, C.byref(q)) == 1
