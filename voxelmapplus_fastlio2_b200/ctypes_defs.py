"""ctypes mirrors of the POD structs in include/vmp_b200.h (field order must match)."""
import ctypes as C

import numpy as np


class VmpConfig(C.Structure):
    """vmp_config == lio::LIOConfig (reference lio_builder.h:17-42) + device additions."""
    _fields_ = [
        ("opti_max_iter", C.c_int),
        ("na", C.c_double), ("ng", C.c_double), ("nba", C.c_double), ("nbg", C.c_double),
        ("imu_init_num", C.c_int),
        ("r_il", C.c_double * 9),
        ("p_il", C.c_double * 3),
        ("gravity_align", C.c_int),
        ("estimate_ext", C.c_int),
        ("scan_resolution", C.c_double),
        ("voxel_size", C.c_double),
        ("update_size_thresh", C.c_int),
        ("max_point_thresh", C.c_int),
        ("plane_thresh", C.c_double),
        ("ranging_cov", C.c_double),
        ("angle_cov", C.c_double),
        ("merge_thresh_for_angle", C.c_double),
        ("merge_thresh_for_distance", C.c_double),
        ("map_capacity", C.c_int),
        ("max_points_per_scan", C.c_int),
        ("device", C.c_int),
    ]


def default_config(**overrides) -> VmpConfig:
    """lio::LIOConfig defaults (lio_builder.h:19-41); scan_resolution=0 keeps pcl::VoxelGrid off the parity path."""
    c = VmpConfig()
    c.opti_max_iter = 5
    c.na, c.ng, c.nba, c.nbg = 0.01, 0.01, 0.0001, 0.0001
    c.imu_init_num = 20
    c.r_il[:] = [1, 0, 0, 0, 1, 0, 0, 0, 1]
    c.p_il[:] = [0, 0, 0]
    c.gravity_align = 1
    c.estimate_ext = 0
    c.scan_resolution = 0.0
    c.voxel_size = 0.5
    c.update_size_thresh = 10
    c.max_point_thresh = 100
    c.plane_thresh = 0.01
    c.ranging_cov = 0.04
    c.angle_cov = 0.1
    c.merge_thresh_for_angle = 0.1
    c.merge_thresh_for_distance = 0.04
    c.map_capacity = 100000
    c.max_points_per_scan = 32768
    c.device = 0
    for k, v in overrides.items():
        if k in ("r_il", "p_il"):
            getattr(c, k)[:] = list(v)
        else:
            if not hasattr(c, k):
                raise AttributeError(k)
            setattr(c, k, v)
    return c


class VmpState(C.Structure):
    """vmp_state == kf::State (reference ieskf.h:31-62)."""
    _fields_ = [
        ("pos", C.c_double * 3), ("rot", C.c_double * 9), ("rot_ext", C.c_double * 9),
        ("pos_ext", C.c_double * 3), ("vel", C.c_double * 3), ("bg", C.c_double * 3),
        ("ba", C.c_double * 3), ("g", C.c_double * 3),
    ]

    def copy(self):
        o = VmpState()
        C.memmove(C.byref(o), C.byref(self), C.sizeof(VmpState))
        return o

    def as_dict(self):
        return {n: np.array(getattr(self, n)[:]) for n, _ in self._fields_}

    @staticmethod
    def identity():
        s = VmpState()
        s.rot[:] = [1, 0, 0, 0, 1, 0, 0, 0, 1]
        s.rot_ext[:] = [1, 0, 0, 0, 1, 0, 0, 0, 1]
        s.g[:] = [0, 0, -9.81]
        return s


class VmpPlane(C.Structure):
    _fields_ = [
        ("key", C.c_int64 * 3), ("mean", C.c_double * 3), ("ppt", C.c_double * 9),
        ("norm", C.c_double * 3), ("cov", C.c_double * 36), ("center", C.c_double * 3),
        ("n", C.c_int32), ("n_temp", C.c_int32), ("newly_add_point", C.c_int32),
        ("flags", C.c_uint32), ("group", C.c_uint64), ("lru_rank", C.c_uint64),
    ]


PLANE_DTYPE = np.dtype([
    ("key", np.int64, 3), ("mean", np.float64, 3), ("ppt", np.float64, 9), ("norm", np.float64, 3),
    ("cov", np.float64, 36), ("center", np.float64, 3), ("n", np.int32), ("n_temp", np.int32),
    ("newly_add_point", np.int32), ("flags", np.uint32), ("group", np.uint64), ("lru_rank", np.uint64),
])
assert PLANE_DTYPE.itemsize == C.sizeof(VmpPlane), (PLANE_DTYPE.itemsize, C.sizeof(VmpPlane))

F_INIT, F_PLANE, F_UPDATE_ENABLE, F_MERGED = 1, 2, 4, 8


class VmpStdVoxel(C.Structure):
    """vmp_std_voxel: one VoxelNode of STDManager::buildVoxels (std_matcher/src/std_manager/descriptor.h:79-95)."""
    _fields_ = [
        ("key", C.c_int64 * 3), ("count", C.c_int32), ("flags", C.c_uint32), ("sum", C.c_double * 3), ("ppt", C.c_double * 9),
        ("mean", C.c_double * 3), ("lamdas", C.c_double * 3), ("norms", C.c_double * 9),
    ]


STD_VOXEL_DTYPE = np.dtype([
    ("key", np.int64, 3), ("count", np.int32), ("flags", np.uint32), ("sum", np.float64, 3), ("ppt", np.float64, 9),
    ("mean", np.float64, 3), ("lamdas", np.float64, 3), ("norms", np.float64, 9),
])
assert STD_VOXEL_DTYPE.itemsize == C.sizeof(VmpStdVoxel), (STD_VOXEL_DTYPE.itemsize, C.sizeof(VmpStdVoxel))
STD_F_VALID, STD_F_PLANE = 1, 2


class VmpUpdateStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "n_points", "n_ins", "n_touch", "n_created", "n_refit", "refit_points", "n_full",
        "n_mergeprobe", "n_merge", "n_evicted", "map_size", "n_mergevox", "n_skipped")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class VmpScanStats(C.Structure):
    _fields_ = [("iters", C.c_int32), ("effect_num", C.c_int32 * 8), ("converged", C.c_int32),
                ("map", VmpUpdateStats), ("gpu_ms", C.c_float), ("host_ms", C.c_float)]


K_COUNT = 25        # VMP_K_COUNT in include/vmp_b200.h


class VmpImu(C.Structure):
    _fields_ = [("acc", C.c_double * 3), ("gyro", C.c_double * 3), ("timestamp", C.c_double)]


IMU_DTYPE = np.dtype([("acc", np.float64, 3), ("gyro", np.float64, 3), ("timestamp", np.float64)])
assert IMU_DTYPE.itemsize == C.sizeof(VmpImu)


def map_update_bytes(st: "VmpUpdateStats | dict") -> int:
    """Algorithmic bytes of one map update (SURVEY.md §8d / DESIGN.md)."""
    d = st if isinstance(st, dict) else st.as_dict()
    return (144 * d["n_ins"] + 160 * d["n_touch"] + 72 * d["refit_points"] + 432 * d["n_refit"]
            + 32 * d["n_full"] + 192 * d["n_mergevox"] + 672 * d["n_merge"])


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))
