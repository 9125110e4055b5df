"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (see oracle/oracle.cpp header).

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from realtime_robot_b200.params import IcpParams, NativeParams, PoseResult, RansacParams, RegisterParams, Surface

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_radius_neighbors.restype = C.c_longlong
        _LIB.orc_harris3d.restype = C.c_int
        _LIB.orc_hypothesis.restype = C.c_int
        _LIB.orc_get_threads.restype = C.c_int
        _LIB.orc_native_pair_score.restype = C.c_float
        _LIB.orc_native_occupancy.restype = C.c_int
        _LIB.orc_native_tdf_voxels.restype = C.c_int
        _LIB.orc_hull_area.restype = C.c_double
        _LIB.orc_plane_areas.restype = C.c_int
        _LIB.orc_plane_segment.restype = C.c_int
        _LIB.orc_plane_class.restype = C.c_int
    return _LIB


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def set_threads(t: int):
    lib().orc_set_threads(C.c_int(t))
    return lib().orc_get_threads()


def tdf(occ, dim=30):
    occ = np.ascontiguousarray(occ, dtype=np.int32).reshape(-1, 3)
    out = np.zeros(dim ** 3, dtype=np.float32)
    lib().orc_tdf(_p(occ, C.c_int), C.c_int(len(occ)), C.c_int(dim), _p(out))
    return out


def radius_neighbors(xyz1, radius, method=1):
    xyz1 = _f(xyz1)
    n = len(xyz1)
    counts = np.zeros(n, dtype=np.int32)
    offsets = np.zeros(n + 1, dtype=np.int64)
    total = lib().orc_radius_neighbors(_p(xyz1), n, C.c_float(radius), method, _p(counts, C.c_int),
                                       _p(offsets, C.c_longlong), None, C.c_longlong(0))
    idx = np.zeros(max(total, 1), dtype=np.int32)
    lib().orc_radius_neighbors(_p(xyz1), n, C.c_float(radius), method, _p(counts, C.c_int),
                               _p(offsets, C.c_longlong), _p(idx, C.c_int), C.c_longlong(total))
    return counts, offsets, idx[:total]


def nearest(tgt, q, method=1):
    tgt, q = _f(tgt), _f(q)
    idx = np.zeros(len(q), dtype=np.int32)
    d2 = np.zeros(len(q), dtype=np.float32)
    lib().orc_nearest(_p(tgt), len(tgt), _p(q), len(q), method, _p(idx, C.c_int), _p(d2))
    return idx, d2


def normals(xyz1, radius, mode=0):
    xyz1 = _f(xyz1)
    out = np.zeros((len(xyz1), 4), dtype=np.float32)
    lib().orc_normals(_p(xyz1), len(xyz1), C.c_float(radius), mode, _p(out))
    return out


def harris3d(xyz1, normals4, radius, threshold, nms=1, refine=1):
    xyz1, normals4 = _f(xyz1), _f(normals4)
    n = len(xyz1)
    resp = np.zeros(n, dtype=np.float32)
    kidx = np.zeros(n, dtype=np.int32)
    kxyz = np.zeros((n, 4), dtype=np.float32)
    m = lib().orc_harris3d(_p(xyz1), n, _p(normals4), C.c_float(radius), C.c_float(threshold), nms, refine,
                           _p(resp), _p(kidx, C.c_int), _p(kxyz), n)
    return resp, kidx[:m].copy(), kxyz[:m].copy()


def fpfh(xyz1, normals4, radius):
    xyz1, normals4 = _f(xyz1), _f(normals4)
    out = np.zeros((len(xyz1), 33), dtype=np.float32)
    lib().orc_fpfh(_p(xyz1), len(xyz1), _p(normals4), C.c_float(radius), _p(out))
    return out


def match_features(fa, fb, k):
    fa, fb = _f(fa), _f(fb)
    idx = np.zeros((len(fa), k), dtype=np.int32)
    dist = np.zeros((len(fa), k), dtype=np.float32)
    lib().orc_match_features(_p(fa), len(fa), _p(fb), len(fb), k, _p(idx, C.c_int), _p(dist))
    return idx, dist


def ransac(src, tgt, knn, params: RansacParams) -> PoseResult:
    src, tgt = _f(src), _f(tgt)
    knn = np.ascontiguousarray(knn, dtype=np.int32)
    res = PoseResult()
    lib().orc_ransac_prerejective(_p(src), len(src), _p(tgt), len(tgt), _p(knn, C.c_int), knn.shape[1],
                                  C.byref(params), C.byref(res))
    return res


def hypothesis(src, tgt, knn, params: RansacParams, h: int):
    src, tgt = _f(src), _f(tgt)
    knn = np.ascontiguousarray(knn, dtype=np.int32)
    s6 = np.zeros(6, dtype=np.int32)
    pose = np.zeros(16, dtype=np.float32)
    ok = lib().orc_hypothesis(_p(src), len(src), _p(tgt), len(tgt), _p(knn, C.c_int), knn.shape[1], C.byref(params),
                              C.c_longlong(h), _p(s6, C.c_int), _p(pose))
    return ok, s6, pose.reshape(4, 4).T.copy()


def icp(src, tgt, params: IcpParams, init=None, tgt_normals=None) -> PoseResult:
    src, tgt = _f(src), _f(tgt)
    res = PoseResult()
    ini = None if init is None else np.ascontiguousarray(np.asarray(init, dtype=np.float32).reshape(4, 4).T).reshape(16)
    if params.estimator == 1:
        nrm = _f(tgt_normals)
        assert nrm.shape == (len(tgt), 4), "estimator 1 (point-to-plane) needs the target's normals"
        lib().orc_icp_normals(_p(src), len(src), _p(tgt), len(tgt), _p(nrm), C.byref(params), _p(ini), C.byref(res))
    else:
        lib().orc_icp(_p(src), len(src), _p(tgt), len(tgt), C.byref(params), _p(ini), C.byref(res))
    return res


def pose_from_pairs(src, tgt):
    src, tgt = _f(src), _f(tgt)
    pose = np.zeros(16, dtype=np.float32)
    lib().orc_pose_from_pairs(_p(src), _p(tgt), len(src), _p(pose))
    return pose.reshape(4, 4).T.copy()


def transform(xyz1, pose):
    xyz1 = _f(xyz1)
    out = np.zeros_like(xyz1)
    m = np.ascontiguousarray(np.asarray(pose, dtype=np.float32).reshape(4, 4).T).reshape(16)
    lib().orc_transform(_p(xyz1), len(xyz1), _p(m), _p(out))
    return out


def jacobi(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = a.shape[0]
    ev = np.zeros(n)
    vec = np.zeros((n, n))
    lib().orc_jacobi(a.ctypes.data_as(C.POINTER(C.c_double)), n, ev.ctypes.data_as(C.POINTER(C.c_double)),
                     vec.ctypes.data_as(C.POINTER(C.c_double)))
    return ev, vec


def horn_top_eigvec(n44):
    """(ok, q): the characteristic-quartic route of horn_pose for a symmetric 4x4."""
    a = np.ascontiguousarray(n44, dtype=np.float64)
    q = np.zeros(4)
    lib().orc_horn_top_eigvec.restype = C.c_int
    ok = lib().orc_horn_top_eigvec(a.ctypes.data_as(C.POINTER(C.c_double)), q.ctypes.data_as(C.POINTER(C.c_double)))
    return bool(ok), q


def register(model, scene, params: RegisterParams) -> PoseResult:
    model, scene = _f(model), _f(scene)
    res = PoseResult()
    lib().orc_register(_p(model), len(model), _p(scene), len(scene), C.byref(params), C.byref(res))
    return res


def register_many(models, scene, params: RegisterParams):
    """rtr_register_many's counterpart: the scan's stages once (RealTimeRobot.cpp:45-60), then every model against it."""
    models = [_f(m) for m in models]
    scene = _f(scene)
    k = len(models)
    ptrs = (C.POINTER(C.c_float) * k)(*[_p(m) for m in models])
    ns = (C.c_int * k)(*[len(m) for m in models])
    res = (PoseResult * k)()
    lib().orc_register_many(ptrs, ns, k, _p(scene), len(scene), C.byref(params), res)
    return [PoseResult.from_buffer_copy(bytes(res[i])) for i in range(k)]


# ------------------------------------------------------------------ reference-native descriptor path (oracle/native.cpp)
def native_keypoint_descriptors(xyz1, kp_xyz1, params: NativeParams, with_tdf=True):
    xyz1, kp = _f(xyz1), _f(kp_xyz1)
    n = len(kp)
    number, count = np.zeros(n, np.int32), np.zeros(n, np.int32)
    tdfs = np.zeros((n, 27000), np.float32)
    vox = np.zeros(n, np.int32)
    occ = []
    idx = np.zeros(max(len(xyz1), 1), np.int32)
    tri = np.zeros((32768, 3), np.int32)
    for k in range(n):
        num = C.c_int()
        c = lib().orc_native_occupancy(_p(xyz1), len(xyz1), _p(kp[k:k + 1]), C.byref(params), _p(idx, C.c_int), len(xyz1), C.byref(num))
        number[k], count[k] = num.value, c
        pts = np.ascontiguousarray(xyz1[idx[:c]])
        occ.append(pts)
        if with_tdf:
            nt = lib().orc_native_tdf_voxels(_p(pts), c, _p(kp[k:k + 1]), C.byref(params), _p(tri, C.c_int), 32768)
            vox[k] = nt
            tdfs[k] = tdf(tri[:nt], int(params.tdf_half / params.resolution * 2))
    return number, count, tdfs, vox, occ


def native_pair_score(model_kp, model_tdf, scan_occ, scan_kp, params: NativeParams):
    mk, sk, occ, t = _f(model_kp).reshape(1, 4), _f(scan_kp).reshape(1, 4), _f(scan_occ).reshape(-1, 4), _f(model_tdf)
    bs = C.c_int()
    T = np.zeros(16, np.float32)
    sc = lib().orc_native_pair_score(_p(mk), _p(t), _p(occ), len(occ), _p(sk), C.byref(params), C.byref(bs), _p(T))
    return float(sc), bs.value, T.reshape(4, 4).T.copy()


def native_register(model, model_kp, scan, scan_kp, params: NativeParams) -> PoseResult:
    model, scan, mk, sk = _f(model), _f(scan), _f(model_kp).reshape(-1, 4), _f(scan_kp).reshape(-1, 4)
    res = PoseResult()
    lib().orc_native_register(_p(model), len(model), _p(mk), len(mk), _p(scan), len(scan), _p(sk), len(sk), C.byref(params), C.byref(res))
    return res


# ------------------------------------------------------------------ getArea (oracle/planes.cpp)
def plane_areas(xyz1, capacity=256):
    xyz1 = _f(xyz1)
    buf = (Surface * capacity)()
    n = lib().orc_plane_areas(_p(xyz1), len(xyz1), buf, capacity)
    return [buf[i] for i in range(min(n, capacity))]


def hull_area(xyz1):
    xyz1 = _f(xyz1)
    dim = C.c_int()
    a = lib().orc_hull_area(_p(xyz1), len(xyz1), C.byref(dim))
    return float(a), dim.value


def plane_segment(xyz1, threshold=0.005, max_iterations=150):
    xyz1 = _f(xyz1)
    coeff = np.zeros(4, np.float32)
    idx = np.zeros(max(len(xyz1), 1), np.int32)
    its = C.c_int()
    m = lib().orc_plane_segment(_p(xyz1), len(xyz1), C.c_float(threshold), max_iterations, _p(coeff), _p(idx, C.c_int), C.byref(its))
    return coeff, idx[:m].copy(), its.value
