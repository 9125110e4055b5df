"""ctypes mirrors of the parameter / result structs in include/rtr.h (shared by the CUDA library binding and,
for struct layout only, by the test oracle's binding)."""
import ctypes as C

import numpy as np


class RansacParams(C.Structure):
    _fields_ = [("max_iterations", C.c_longlong), ("hypothesis_begin", C.c_longlong), ("hypothesis_end", C.c_longlong),
                ("seed", C.c_ulonglong), ("correspondence_k", C.c_int), ("similarity_threshold", C.c_float),
                ("max_correspondence_distance", C.c_float), ("inlier_fraction", C.c_float)]


class IcpParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("force_iterations", C.c_int), ("max_correspondence_distance", C.c_float),
                ("estimator", C.c_int), ("mse_threshold_absolute", C.c_double)]


class RegisterParams(C.Structure):
    _fields_ = [("normal_radius", C.c_float), ("harris_radius", C.c_float), ("harris_threshold", C.c_float),
                ("harris_nms", C.c_int), ("harris_refine", C.c_int), ("fpfh_radius", C.c_float), ("run_icp", C.c_int),
                ("pad_", C.c_int), ("ransac", RansacParams), ("icp", IcpParams)]


class PoseResult(C.Structure):
    _fields_ = [("pose", C.c_float * 16), ("fitness", C.c_float), ("inliers", C.c_int), ("hypothesis", C.c_longlong),
                ("evaluated", C.c_longlong), ("converged", C.c_int), ("iterations", C.c_int), ("model_id", C.c_int),
                ("n_keypoints_src", C.c_int), ("n_keypoints_tgt", C.c_int), ("pad_", C.c_int * 5)]

    def matrix(self) -> np.ndarray:
        """4x4 float32, row-indexed [r, c] (the struct stores Eigen's column-major order)."""
        return np.array(self.pose, dtype=np.float32).reshape(4, 4).T.copy()

    def as_dict(self):
        return {"pose": self.matrix().tolist(), "fitness": float(self.fitness), "inliers": int(self.inliers),
                "hypothesis": int(self.hypothesis), "evaluated": int(self.evaluated), "converged": int(self.converged),
                "iterations": int(self.iterations), "model_id": int(self.model_id),
                "n_keypoints_src": int(self.n_keypoints_src), "n_keypoints_tgt": int(self.n_keypoints_tgt)}


assert C.sizeof(PoseResult) == 128, C.sizeof(PoseResult)


class Surface(C.Structure):
    """rtr_surface: one peeled plane of ModelPoint::getArea (Surface of key_point.h:47-51 plus bookkeeping)."""
    _fields_ = [("area", C.c_double), ("coefficients", C.c_float * 4), ("is_vertical", C.c_int), ("inliers", C.c_int),
                ("dimension", C.c_int), ("iterations", C.c_int), ("kept", C.c_int), ("pad_", C.c_int)]

    def as_tuple(self):
        return (float(self.area), tuple(float(v) for v in self.coefficients), int(self.is_vertical), int(self.inliers),
                int(self.dimension), int(self.iterations), int(self.kept))


class NativeParams(C.Structure):
    """rtr_native_params: the reference's own descriptor path (key_point.h / matching.h / function.h literals)."""
    _fields_ = [("resolution", C.c_float), ("occ_half", C.c_float), ("tdf_half", C.c_float), ("pair_gate", C.c_float),
                ("consensus_distance", C.c_float), ("consensus_score", C.c_float), ("quirk_skip_first_voxel", C.c_int),
                ("quirk_running_score", C.c_int), ("quirk_integer_screens", C.c_int), ("use_plane_areas", C.c_int)]


def default_native_params() -> NativeParams:
    p = NativeParams()
    p.resolution, p.occ_half, p.tdf_half = 0.01, 0.1, 0.15
    p.pair_gate, p.consensus_distance, p.consensus_score = 3.0, 0.15, 100.0
    p.use_plane_areas = 1
    return p


def default_register_params() -> RegisterParams:
    """Defaults of include/rtr.h (reference literals where the reference has the stage, PCL-tutorial values otherwise)."""
    p = RegisterParams()
    p.normal_radius = 0.05
    p.harris_radius = 0.05
    p.harris_threshold = 0.01
    p.harris_nms = 1
    p.harris_refine = 1
    p.fpfh_radius = 0.10
    p.run_icp = 1
    p.ransac.max_iterations = 50000
    p.ransac.hypothesis_begin = 0
    p.ransac.hypothesis_end = 0
    p.ransac.seed = 20170427
    p.ransac.correspondence_k = 5
    p.ransac.similarity_threshold = 0.9
    p.ransac.max_correspondence_distance = 0.0365
    p.ransac.inlier_fraction = 0.25
    p.icp.max_iterations = 10
    p.icp.force_iterations = 0
    p.icp.max_correspondence_distance = 0.0
    p.icp.mse_threshold_absolute = 1e-12
    return p


def pose_to_colmajor(m) -> np.ndarray:
    """4x4 [r, c] array -> 16 floats in the C ABI's column-major order."""
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).reshape(4, 4).T).reshape(16)
