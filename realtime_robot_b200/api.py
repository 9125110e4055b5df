"""Host-side mirror of the reference's registration surface over the C ABI (include/rtr.h).

Names follow the reference: a cloud object owns the points and its per-cloud stages the way ModelPoint / ScanPoint do
(model_point.h:81-96, scan_point.h:43-54: getKeypoint), and the free functions mirror function.h (keyPointICP -> icp,
Ransac -> ransac_prerejective).  Everything here is plumbing: ctypes calls into librtr.so.  No compute happens in Python
and there is no CPU fallback.
"""
import atexit
import ctypes as C
import os
import weakref

import numpy as np

from . import _lib
from .params import (IcpParams, NativeParams, PoseResult, RansacParams, RegisterParams, Surface, default_native_params,  # noqa: F401
                     default_register_params, pose_to_colmajor)


def _f32(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if cols is not None and (a.ndim != 2 or a.shape[1] != cols):
        raise ValueError(f"expected an (N, {cols}) float32 array, got {a.shape}")
    return a


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


_LIVE = weakref.WeakSet()       # contexts and clouds still holding device resources


@atexit.register
def _release_all():
    """Free clouds, then contexts, while the CUDA runtime is still loaded (destructors that run during interpreter teardown
    would call into a runtime that has already shut down)."""
    live = list(_LIVE)
    for o in live:
        if isinstance(o, Cloud):
            o.free()
    for o in live:
        if isinstance(o, Context):
            o.close()


class Context:
    """One per GPU: device id, one stream, stream-ordered memory pool (rtr_context_create)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _lib.check("rtr_context_create", _lib.lib().rtr_context_create(device, C.byref(self._h)))
        self.device = device
        _LIVE.add(self)

    def sync(self):
        _lib.check("rtr_context_sync", _lib.lib().rtr_context_sync(self._h))

    @property
    def launches(self) -> int:
        return int(_lib.lib().rtr_context_launches(self._h))

    def record(self, slot: int):
        _lib.check("rtr_event_record", _lib.lib().rtr_event_record(self._h, slot))

    def elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        _lib.check("rtr_event_elapsed_ms", _lib.lib().rtr_event_elapsed_ms(self._h, a, b, C.byref(ms)))
        return float(ms.value)

    def profile_begin(self):
        _lib.check("rtr_profile_begin", _lib.lib().rtr_profile_begin(self._h))

    def profile_end(self) -> dict:
        """{tag: (count, total_ms)} for every stream operation since profile_begin."""
        buf = C.create_string_buffer(1 << 16)
        _lib.check("rtr_profile_end", _lib.lib().rtr_profile_end(self._h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            tag, cnt, ms = line.rsplit(" ", 2)
            out[tag] = (int(cnt), float(ms))
        return out

    def close(self):
        if self._h:
            _lib.lib().rtr_context_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


PCD_ASCII, PCD_BINARY, PCD_BINARY_COMPRESSED = 0, 1, 2


def pcd_info(path: str):
    """(number of points, DATA mode) of a PCD v0.7 file."""
    n, m = C.c_int(), C.c_int()
    _lib.check("rtr_pcd_info", _lib.lib().rtr_pcd_info(os.fsencode(path), C.byref(n), C.byref(m)))
    return n.value, m.value


def read_pcd(path: str) -> np.ndarray:
    """pcl::io::loadPCDFile into host memory: (n, 4) float32 (x, y, z, 1); ascii, binary and binary_compressed."""
    n, _ = pcd_info(path)
    out = np.zeros((n, 4), dtype=np.float32)
    got = C.c_int()
    _lib.check("rtr_pcd_read", _lib.lib().rtr_pcd_read(os.fsencode(path), _ptr(out), n, C.byref(got)))
    return out


def write_pcd(path: str, xyz1, mode: int = PCD_ASCII):
    """pcl::io::savePCDFileASCII (mode 0) / savePCDFileBinary (1) / savePCDFileBinaryCompressed (2)."""
    xyz1 = _f32(xyz1, 4)
    _lib.check("rtr_pcd_write", _lib.lib().rtr_pcd_write(os.fsencode(path), _ptr(xyz1), len(xyz1), mode))


class Cloud:
    """Device-resident pcl::PointCloud<pcl::PointXYZ> (n x 16 B) plus its cached stages."""

    def __init__(self, ctx: Context, xyz1=None, device_ptr=None, n=None):
        self.ctx = ctx
        self._h = C.c_void_p()
        if device_ptr is not None:
            _lib.check("rtr_cloud_from_device", _lib.lib().rtr_cloud_from_device(ctx._h, C.c_void_p(device_ptr), n, C.byref(self._h)))
            self.n = n
        else:
            xyz1 = _f32(xyz1, 4)
            self.n = len(xyz1)
            _lib.check("rtr_cloud_upload", _lib.lib().rtr_cloud_upload(ctx._h, _ptr(xyz1), self.n, C.byref(self._h)))
            ctx.sync()   # the numpy buffer may be released as soon as we return
        _LIVE.add(self)

    def free(self):
        if self._h:
            _lib.lib().rtr_cloud_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @classmethod
    def from_pcd(cls, ctx: Context, path: str) -> "Cloud":
        """pcl::io::loadPCDFile straight to the device: file -> pinned staging -> one async copy (rtr_pcd_load)."""
        self = cls.__new__(cls)
        self.ctx = ctx
        self._h = C.c_void_p()
        _lib.check("rtr_pcd_load", _lib.lib().rtr_pcd_load(ctx._h, os.fsencode(path), C.byref(self._h)))
        self.n = int(_lib.lib().rtr_cloud_size(self._h))
        _LIVE.add(self)
        return self

    def save_pcd(self, path: str, mode: int = PCD_ASCII):
        """pcl::io::savePCDFileASCII (mode 0) / binary (1) / binary_compressed (2) of the device cloud."""
        _lib.check("rtr_cloud_save", _lib.lib().rtr_cloud_save(self._h, os.fsencode(path), mode))

    def prepare(self, params: RegisterParams) -> None:
        """The offline half (RealTimeRobot.cpp:124-165): normals, Harris corners and FPFH rows with the stage parameters of
        `params`, kept on the cloud until reset() / free() — what register_prepared expects of every model."""
        _lib.check("rtr_cloud_prepare", _lib.lib().rtr_cloud_prepare(self._h, C.byref(params)))

    def reset(self):
        """Forget every cached stage (the points stay resident)."""
        _lib.check("rtr_cloud_reset", _lib.lib().rtr_cloud_reset(self._h))

    def download(self) -> np.ndarray:
        out = np.zeros((self.n, 4), dtype=np.float32)
        _lib.check("rtr_cloud_download", _lib.lib().rtr_cloud_download(self._h, _ptr(out)))
        return out

    def transform(self, pose):
        m = pose_to_colmajor(pose)
        _lib.check("rtr_cloud_transform", _lib.lib().rtr_cloud_transform(self._h, _ptr(m)))

    def radius_neighbors(self, radius: float, counts_only: bool = False):
        counts = np.zeros(self.n, dtype=np.int32)
        offsets = np.zeros(self.n + 1, dtype=np.int64)
        total = C.c_longlong()
        _lib.check("rtr_radius_neighbors", _lib.lib().rtr_radius_neighbors(self._h, radius, _ptr(counts), _ptr(offsets), None, 0, C.byref(total)))
        if counts_only:
            return counts, offsets, None
        idx = np.zeros(max(total.value, 1), dtype=np.int32)
        _lib.check("rtr_radius_neighbors", _lib.lib().rtr_radius_neighbors(self._h, radius, _ptr(counts), _ptr(offsets), _ptr(idx), total.value, C.byref(total)))
        return counts, offsets, idx[:total.value]

    def nearest(self, queries_xyz1):
        q = _f32(queries_xyz1, 4)
        idx = np.zeros(len(q), dtype=np.int32)
        d2 = np.zeros(len(q), dtype=np.float32)
        _lib.check("rtr_nearest", _lib.lib().rtr_nearest(self._h, _ptr(q), len(q), _ptr(idx), _ptr(d2)))
        return idx, d2

    def normals(self, radius: float, mode: int = 0) -> np.ndarray:
        """pcl::NormalEstimation; mode 0: exact (fp64, the parity mode), mode 1: PCL-float-faithful (rtr_normals_mode)."""
        out = np.zeros((self.n, 4), dtype=np.float32)
        if mode == 0:
            _lib.check("rtr_normals", _lib.lib().rtr_normals(self._h, radius, _ptr(out)))
        else:
            _lib.check("rtr_normals_mode", _lib.lib().rtr_normals_mode(self._h, radius, mode, _ptr(out)))
        return out

    def harris3d(self, radius: float, threshold: float, nms: int = 1, refine: int = 1):
        """ModelPoint::getKeypoint / ScanPoint::getKeypoint: returns (response, corner indices, corner xyz1)."""
        resp = np.zeros(self.n, dtype=np.float32)
        idx = np.zeros(max(self.n, 1), dtype=np.int32)
        xyz = np.zeros((max(self.n, 1), 4), dtype=np.float32)
        m = C.c_int()
        _lib.check("rtr_harris3d", _lib.lib().rtr_harris3d(self._h, radius, threshold, nms, refine, _ptr(resp), _ptr(idx), _ptr(xyz), self.n, C.byref(m)))
        return resp, idx[:m.value].copy(), xyz[:m.value].copy()

    def fpfh(self, radius: float) -> np.ndarray:
        out = np.zeros((self.n, 33), dtype=np.float32)
        _lib.check("rtr_fpfh", _lib.lib().rtr_fpfh(self._h, radius, _ptr(out)))
        return out

    def fpfh_at(self, radius: float, query_index) -> np.ndarray:
        """FPFH at a subset of the points (PCL: setInputCloud(keypoints) + setSearchSurface(cloud)); rows follow query_index."""
        q = np.ascontiguousarray(query_index, dtype=np.int32)
        out = np.zeros((len(q), 33), dtype=np.float32)
        _lib.check("rtr_fpfh_at", _lib.lib().rtr_fpfh_at(self._h, radius, _ptr(q), len(q), _ptr(out)))
        return out

    def match_features(self, target: "Cloud", k: int):
        idx = np.zeros((self.n, k), dtype=np.int32)
        dist = np.zeros((self.n, k), dtype=np.float32)
        _lib.check("rtr_match_features", _lib.lib().rtr_match_features(self._h, target._h, k, _ptr(idx), _ptr(dist)))
        return idx, dist


def ransac_prerejective(source: Cloud, target: Cloud, params: RansacParams) -> PoseResult:
    res = PoseResult()
    _lib.check("rtr_ransac_prerejective", _lib.lib().rtr_ransac_prerejective(source._h, target._h, C.byref(params), C.byref(res)))
    return res


def icp(source: Cloud, target: Cloud, params: IcpParams, init=None) -> PoseResult:
    """keyPointICP (function.h:111-123): source = model, target = scan."""
    res = PoseResult()
    m = None if init is None else pose_to_colmajor(init)
    _lib.check("rtr_icp", _lib.lib().rtr_icp(source._h, target._h, C.byref(params), _ptr(m), C.byref(res)))
    return res


def register(model: Cloud, scene: Cloud, params: RegisterParams) -> PoseResult:
    res = PoseResult()
    _lib.check("rtr_register", _lib.lib().rtr_register(model._h, scene._h, C.byref(params), C.byref(res)))
    return res


def register_host(ctx: Context, model_xyz1, scene_xyz1, params: RegisterParams) -> PoseResult:
    """Host clouds in, host result out: upload + register + free inside one C-ABI call."""
    m, s = _f32(model_xyz1, 4), _f32(scene_xyz1, 4)
    res = PoseResult()
    _lib.check("rtr_register_host", _lib.lib().rtr_register_host(ctx._h, _ptr(m), len(m), _ptr(s), len(s), C.byref(params), C.byref(res)))
    return res


def register_begin(model: Cloud, scene: Cloud, params: RegisterParams) -> None:
    """Queue a whole registration on the clouds' context and return (no host synchronisation); see register_end."""
    _lib.check("rtr_register_begin", _lib.lib().rtr_register_begin(model._h, scene._h, C.byref(params)))


def register_host_begin(ctx: Context, model_xyz1, scene_xyz1, params: RegisterParams) -> None:
    """Same from host clouds (pinned memory keeps the uploads asynchronous).  The arrays must stay alive until register_end."""
    m, s = _f32(model_xyz1, 4), _f32(scene_xyz1, 4)
    ctx._inflight = (m, s)
    _lib.check("rtr_register_host_begin", _lib.lib().rtr_register_host_begin(ctx._h, _ptr(m), len(m), _ptr(s), len(s), C.byref(params)))


def register_end(ctx: Context) -> PoseResult:
    """Wait for the registration in flight on `ctx` and return its record."""
    res = PoseResult()
    _lib.check("rtr_register_end", _lib.lib().rtr_register_end(ctx._h, C.byref(res)))
    ctx._inflight = None
    return res


def _records(buf, n):
    return [PoseResult.from_buffer_copy(bytes(buf[i])) for i in range(n)]


def register_many(models, scene: Cloud, params: RegisterParams):
    """One scan against many database models (RealTimeRobot.cpp:45-104 once per model, README.md:10): the scan's stages run
    once, the models' stages share their launches.  Returns one PoseResult per model, equal to register(model, scene)."""
    k = len(models)
    handles = (C.c_void_p * k)(*[m._h for m in models])
    res = (PoseResult * k)()
    _lib.check("rtr_register_many", _lib.lib().rtr_register_many(handles, k, scene._h, C.byref(params), res))
    return _records(res, k)


def register_prepared(models, scene: Cloud, params: RegisterParams):
    """The online half of register_many: every model went through Cloud.prepare(params); the scan's stages run here, once.
    Same records as register_many / register, bit for bit."""
    k = len(models)
    handles = (C.c_void_p * k)(*[m._h for m in models])
    res = (PoseResult * k)()
    _lib.check("rtr_register_prepared", _lib.lib().rtr_register_prepared(handles, k, scene._h, C.byref(params), res))
    return _records(res, k)


def register_prepared_begin(models, scene: Cloud, params: RegisterParams) -> None:
    """Enqueue only (at most 31 models); register_many_end(ctx) waits and returns the records."""
    k = len(models)
    handles = (C.c_void_p * k)(*[m._h for m in models])
    scene.ctx._inflight_many = k
    _lib.check("rtr_register_prepared_begin", _lib.lib().rtr_register_prepared_begin(handles, k, scene._h, C.byref(params)))


def _host_batch(models_xyz1, scene_xyz1):
    ms = [_f32(m, 4) for m in models_xyz1]
    s = _f32(scene_xyz1, 4)
    k = len(ms)
    ptrs = (C.c_void_p * k)(*[m.ctypes.data for m in ms])
    ns = (C.c_int * k)(*[len(m) for m in ms])
    return ms, s, ptrs, ns, k


def register_many_host(ctx: Context, models_xyz1, scene_xyz1, params: RegisterParams):
    """Host clouds in, host records out: uploads + batch + free inside one C-ABI call."""
    ms, s, ptrs, ns, k = _host_batch(models_xyz1, scene_xyz1)
    res = (PoseResult * k)()
    _lib.check("rtr_register_many_host", _lib.lib().rtr_register_many_host(ctx._h, ptrs, ns, k, _ptr(s), len(s), C.byref(params), res))
    return _records(res, k)


def register_many_begin(models, scene: Cloud, params: RegisterParams) -> None:
    k = len(models)
    handles = (C.c_void_p * k)(*[m._h for m in models])
    scene.ctx._inflight_many = k
    _lib.check("rtr_register_many_begin", _lib.lib().rtr_register_many_begin(handles, k, scene._h, C.byref(params)))


def register_many_host_begin(ctx: Context, models_xyz1, scene_xyz1, params: RegisterParams) -> None:
    """The arrays must stay alive until register_many_end (pinned memory keeps the uploads asynchronous)."""
    ms, s, ptrs, ns, k = _host_batch(models_xyz1, scene_xyz1)
    ctx._inflight = (ms, s, ptrs, ns)
    ctx._inflight_many = k
    _lib.check("rtr_register_many_host_begin", _lib.lib().rtr_register_many_host_begin(ctx._h, ptrs, ns, k, _ptr(s), len(s), C.byref(params)))


def register_many_end(ctx: Context):
    k = ctx._inflight_many
    res = (PoseResult * k)()
    _lib.check("rtr_register_many_end", _lib.lib().rtr_register_many_end(ctx._h, res, k))
    ctx._inflight = None
    return _records(res, k)


def register_many_keypoints(ctx: Context, member: int, capacity: int = 64):
    """Refined Harris corners (xyz1) of one member of the last batch: 0..k-1 the models, k the scan; (corners, full count)."""
    out = np.zeros((capacity, 4), dtype=np.float32)
    n = C.c_int()
    rc = _lib.lib().rtr_register_many_keypoints(ctx._h, member, _ptr(out), capacity, C.byref(n))
    if rc not in (0, 3):
        _lib.check("rtr_register_many_keypoints", rc)
    return out[:min(n.value, capacity)].copy(), n.value


def compute_tdf_with_cuda(voxel_grid_occ, voxel_grid_tdf, voxel_grid_dim: int, num_occ: int) -> int:
    """The reference FFI, unchanged (key_point.h:35-36).  voxel_grid_tdf: 27000 float32, modified in place."""
    occ = np.ascontiguousarray(voxel_grid_occ, dtype=np.int32)
    if voxel_grid_tdf.dtype != np.float32 or not voxel_grid_tdf.flags["C_CONTIGUOUS"] or voxel_grid_tdf.size < 27000:
        raise ValueError("voxel_grid_tdf must be a contiguous float32 array of >= 27000 elements")
    return int(_lib.lib().ComputeTDFWithCuda(_ptr(occ) if occ.size else None, _ptr(voxel_grid_tdf), voxel_grid_dim, num_occ))


def tdf_batch(ctx: Context, occ_lists, dim: int = 30) -> np.ndarray:
    """All keypoints of a cloud in one launch; occ_lists: sequence of (k_i, 3) int arrays."""
    offs = np.zeros(len(occ_lists) + 1, dtype=np.int32)
    for i, o in enumerate(occ_lists):
        offs[i + 1] = offs[i] + len(o)
    occ = (np.concatenate([np.asarray(o, dtype=np.int32).reshape(-1, 3) for o in occ_lists]) if len(occ_lists)
           else np.zeros((0, 3), dtype=np.int32))
    occ = np.ascontiguousarray(occ, dtype=np.int32)
    out = np.zeros((len(occ_lists), dim ** 3), dtype=np.float32)
    _lib.check("rtr_tdf_batch", _lib.lib().rtr_tdf_batch(ctx._h, _ptr(occ) if occ.size else None, _ptr(offs), len(occ_lists), dim, _ptr(out)))
    return out


def match_raw(ctx: Context, fa, fb, k: int, reps: int = 1):
    """Feature k-NN on two bare (n, 33) feature arrays (bench helper): uploads them as the FPFH of two placeholder
    clouds is not possible through the cloud API, so this goes through rtr_match_features_raw.  Returns (best ms, stats)."""
    fa, fb = _f32(fa, 33), _f32(fb, 33)
    idx = np.zeros((len(fa), k), dtype=np.int32)
    dist = np.zeros((len(fa), k), dtype=np.float32)
    ms = C.c_float()
    best = None
    for _ in range(max(reps, 1)):
        _lib.check("rtr_match_features_raw", _lib.lib().rtr_match_features_raw(ctx._h, _ptr(fa), len(fa), _ptr(fb), len(fb), k, _ptr(idx), _ptr(dist), C.byref(ms)))
        best = ms.value if best is None else min(best, ms.value)
    st = (C.c_int * 3)()
    _lib.lib().rtr_match_last_stats(ctx._h, st)
    return best, {"idx": idx, "dist": dist, "redo_rows": int(st[0]), "splits": int(st[1]), "observed_err_over_norms": st[2] * 1e-9}


# ------------------------------------------------------------------ the reference's own descriptor path
def native_keypoint_descriptors(cloud: Cloud, kp_xyz1, params: NativeParams, with_tdf: bool = True):
    """KeyPoint::getOccupiedGrid + KeyPoint::get_TSDF for all keypoints: (Number, points in box, TDF [n,27000], voxels)."""
    kp = _f32(kp_xyz1, 4)
    n = len(kp)
    number = np.zeros(n, dtype=np.int32)
    count = np.zeros(n, dtype=np.int32)
    tdf = np.zeros((n, 27000), dtype=np.float32) if with_tdf else None
    vox = np.zeros(n, dtype=np.int32) if with_tdf else None
    _lib.check("rtr_native_keypoint_descriptors", _lib.lib().rtr_native_keypoint_descriptors(
        cloud._h, _ptr(kp) if n else None, n, C.byref(params), _ptr(number), _ptr(count), _ptr(tdf), _ptr(vox)))
    return number, count, tdf, vox


def native_pair_scores(model: Cloud, model_kp, scan: Cloud, scan_kp, params: NativeParams):
    """get_Distance for all pairs: (score [km,ks], best step [km,ks], transforms [km,ks,4,4] row-indexed)."""
    mk, sk = _f32(model_kp, 4), _f32(scan_kp, 4)
    km, ks = len(mk), len(sk)
    score = np.zeros((km, ks), dtype=np.float32)
    best = np.zeros((km, ks), dtype=np.int32)
    tr = np.zeros((km, ks, 16), dtype=np.float32)
    _lib.check("rtr_native_pair_scores", _lib.lib().rtr_native_pair_scores(
        model._h, _ptr(mk) if km else None, km, scan._h, _ptr(sk) if ks else None, ks, C.byref(params), _ptr(score), _ptr(best), _ptr(tr)))
    return score, best, tr.reshape(km, ks, 4, 4).transpose(0, 1, 3, 2).copy()


def native_register(model: Cloud, scan: Cloud, params: NativeParams) -> PoseResult:
    """main() of the reference as intended: Harris -> occupancy / TDF -> yaw sweep -> screens -> exhaustive consensus."""
    res = PoseResult()
    _lib.check("rtr_native_register", _lib.lib().rtr_native_register(model._h, scan._h, C.byref(params), C.byref(res)))
    return res


def plane_areas(cloud: Cloud, capacity: int = 256):
    """ModelPoint::getArea / ScanPoint::get_Area: every peeled plane, in order (Surface records; .kept marks `surface` entries)."""
    buf = (Surface * capacity)()
    n = C.c_int()
    _lib.check("rtr_plane_areas", _lib.lib().rtr_plane_areas(cloud._h, buf, capacity, C.byref(n)))
    return [buf[i] for i in range(n.value)]
