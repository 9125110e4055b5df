"""Loader for librtr.so (the hand-written sm_100a CUDA library behind include/rtr.h).

There is no CPU fallback: if the library is missing or cannot be loaded, importing the compute API raises.
"""
import ctypes as C
import os

from .params import IcpParams, NativeParams, PoseResult, RansacParams, RegisterParams, Surface

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librtr.so")
_LIB = None

# every symbol include/rtr.h declares (tests check that the library exports each one)
EXPORTS = [
    "rtr_default_register_params", "rtr_context_create", "rtr_context_destroy", "rtr_context_sync", "rtr_context_stream",
    "rtr_context_launches", "rtr_profile_begin", "rtr_profile_end", "rtr_cloud_reset", "rtr_event_record", "rtr_event_elapsed_ms", "rtr_cloud_upload", "rtr_cloud_from_device",
    "rtr_cloud_free", "rtr_cloud_size", "rtr_cloud_transform", "rtr_cloud_download", "rtr_radius_neighbors", "rtr_nearest",
    "rtr_normals", "rtr_harris3d", "rtr_fpfh", "rtr_fpfh_at", "rtr_match_features", "rtr_match_features_raw", "rtr_match_last_stats", "rtr_ransac_prerejective", "rtr_icp", "rtr_register",
    "rtr_register_host", "ComputeTDFWithCuda", "rtr_tdf_batch", "rtr_tdf_batch_dev", "rtr_native_default_params", "rtr_native_keypoint_descriptors",
    "rtr_native_pair_scores", "rtr_native_register", "rtr_plane_areas", "rtr_pcd_info", "rtr_pcd_read", "rtr_pcd_load", "rtr_pcd_write",
    "rtr_cloud_save", "rtr_register_begin", "rtr_register_host_begin", "rtr_register_end", "rtr_context_create_prio",
    "rtr_register_many", "rtr_register_many_host", "rtr_register_many_begin", "rtr_register_many_host_begin", "rtr_register_many_end",
    "rtr_cloud_prepare", "rtr_register_prepared", "rtr_register_prepared_begin",
    "rtr_register_many_keypoints", "rtr_comm_unique_id", "rtr_comm_init", "rtr_comm_destroy", "rtr_comm_world", "rtr_allgather_results",
    "rtr_select_best_hypothesis", "rtr_normals_mode", "rtr_comm_gather_batches", "rtr_gathered_results",
]


class RtrError(RuntimeError):
    def __init__(self, fn, code):
        super().__init__(f"{fn} failed with status {code}")
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C realtime_robot_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, ip, fp, ll = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_longlong
        L.rtr_context_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.rtr_context_create_prio.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
        L.rtr_context_destroy.argtypes = [vp]
        L.rtr_context_sync.argtypes = [vp]
        L.rtr_context_stream.argtypes = [vp]
        L.rtr_context_stream.restype = vp
        L.rtr_context_launches.argtypes = [vp]
        L.rtr_context_launches.restype = ll
        L.rtr_event_record.argtypes = [vp, C.c_int]
        L.rtr_event_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, fp]
        L.rtr_profile_begin.argtypes = [vp]
        L.rtr_profile_end.argtypes = [vp, C.c_char_p, C.c_int]
        L.rtr_cloud_reset.argtypes = [vp]
        L.rtr_cloud_upload.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
        L.rtr_cloud_from_device.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
        L.rtr_cloud_free.argtypes = [vp]
        L.rtr_cloud_size.argtypes = [vp]
        L.rtr_cloud_transform.argtypes = [vp, vp]
        L.rtr_cloud_download.argtypes = [vp, vp]
        L.rtr_radius_neighbors.argtypes = [vp, C.c_float, vp, vp, vp, ll, C.POINTER(ll)]
        L.rtr_nearest.argtypes = [vp, vp, C.c_int, vp, vp]
        L.rtr_normals.argtypes = [vp, C.c_float, vp]
        L.rtr_normals_mode.argtypes = [vp, C.c_float, C.c_int, vp]
        L.rtr_harris3d.argtypes = [vp, C.c_float, C.c_float, C.c_int, C.c_int, vp, vp, vp, C.c_int, ip]
        L.rtr_fpfh.argtypes = [vp, C.c_float, vp]
        L.rtr_fpfh_at.argtypes = [vp, C.c_float, vp, C.c_int, vp]
        L.rtr_match_features.argtypes = [vp, vp, C.c_int, vp, vp]
        L.rtr_match_features_raw.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, fp]
        L.rtr_match_last_stats.argtypes = [vp, ip]
        L.rtr_ransac_prerejective.argtypes = [vp, vp, C.POINTER(RansacParams), C.POINTER(PoseResult)]
        L.rtr_icp.argtypes = [vp, vp, C.POINTER(IcpParams), vp, C.POINTER(PoseResult)]
        L.rtr_register.argtypes = [vp, vp, C.POINTER(RegisterParams), C.POINTER(PoseResult)]
        L.rtr_register_host.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.POINTER(RegisterParams), C.POINTER(PoseResult)]
        L.ComputeTDFWithCuda.argtypes = [vp, vp, C.c_int, C.c_int]
        L.rtr_tdf_batch.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp]
        L.rtr_tdf_batch_dev.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp]
        L.rtr_default_register_params.argtypes = [C.POINTER(RegisterParams)]
        L.rtr_native_default_params.argtypes = [C.POINTER(NativeParams)]
        L.rtr_native_keypoint_descriptors.argtypes = [vp, vp, C.c_int, C.POINTER(NativeParams), vp, vp, vp, vp]
        L.rtr_native_pair_scores.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, C.POINTER(NativeParams), vp, vp, vp]
        L.rtr_plane_areas.argtypes = [vp, C.POINTER(Surface), C.c_int, ip]
        L.rtr_native_register.argtypes = [vp, vp, C.POINTER(NativeParams), C.POINTER(PoseResult)]
        L.rtr_register_begin.argtypes = [vp, vp, C.POINTER(RegisterParams)]
        L.rtr_register_host_begin.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.POINTER(RegisterParams)]
        L.rtr_register_end.argtypes = [vp, C.POINTER(PoseResult)]
        L.rtr_register_many.argtypes = [C.POINTER(vp), C.c_int, vp, C.POINTER(RegisterParams), C.POINTER(PoseResult)]
        L.rtr_register_many_host.argtypes = [vp, C.POINTER(vp), ip, C.c_int, vp, C.c_int, C.POINTER(RegisterParams), C.POINTER(PoseResult)]
        L.rtr_register_many_begin.argtypes = [C.POINTER(vp), C.c_int, vp, C.POINTER(RegisterParams)]
        L.rtr_cloud_prepare.argtypes = [vp, C.POINTER(RegisterParams)]
        L.rtr_register_prepared.argtypes = [C.POINTER(vp), C.c_int, vp, C.POINTER(RegisterParams), C.POINTER(PoseResult)]
        L.rtr_register_prepared_begin.argtypes = [C.POINTER(vp), C.c_int, vp, C.POINTER(RegisterParams)]
        L.rtr_register_many_host_begin.argtypes = [vp, C.POINTER(vp), ip, C.c_int, vp, C.c_int, C.POINTER(RegisterParams)]
        L.rtr_register_many_end.argtypes = [vp, C.POINTER(PoseResult), C.c_int]
        L.rtr_register_many_keypoints.argtypes = [vp, C.c_int, vp, C.c_int, ip]
        L.rtr_comm_unique_id.argtypes = [C.c_char_p]
        L.rtr_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
        L.rtr_comm_destroy.argtypes = [vp]
        L.rtr_comm_world.argtypes = [vp, ip, ip]
        L.rtr_comm_gather_batches.argtypes = [vp, C.c_int, C.c_int]
        L.rtr_gathered_results.argtypes = [vp, C.POINTER(PoseResult), C.c_int, ip]
        L.rtr_allgather_results.argtypes = [vp, C.POINTER(PoseResult), C.c_int, C.POINTER(PoseResult)]
        L.rtr_select_best_hypothesis.argtypes = [C.POINTER(PoseResult), C.c_int, C.POINTER(PoseResult)]
        L.rtr_pcd_info.argtypes = [C.c_char_p, ip, ip]
        L.rtr_pcd_read.argtypes = [C.c_char_p, vp, C.c_int, ip]
        L.rtr_pcd_load.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
        L.rtr_pcd_write.argtypes = [C.c_char_p, vp, C.c_int, C.c_int]
        L.rtr_cloud_save.argtypes = [vp, C.c_char_p, C.c_int]
        _LIB = L
    return _LIB


def check(fn, code):
    if code != 0:
        raise RtrError(fn, code)
