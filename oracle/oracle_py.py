"""TEST INFRASTRUCTURE ONLY — Python access to the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package never does.

PARITY UNPINNED: the reference has no tests or golden vectors and cannot be compiled in
this image (Eigen/Sophus/PCL/ROS absent), so this oracle is a from-scratch restatement;
see oracle/README.md.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from voxelmapplus_fastlio2_b200.bindings import HotPath
from voxelmapplus_fastlio2_b200.ctypes_defs import (VmpImu, VmpScanStats, VmpState, VmpUpdateStats, dptr, fptr)

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_DIR, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (reference flags + -ffp-contract=off)."""
    srcs = [os.path.join(_DIR, f) for f in ("oracle.cpp", "oracle_capi.cpp", "oracle.h", "oracle_math.h")]
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _DIR, "-B", "liboracle.so"], check=True, capture_output=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        _lib.orc_eig3.argtypes = [dp, dp, dp]
        _lib.orc_inverse23.argtypes = [dp, dp]
        _lib.orc_so3_exp.argtypes = [dp, dp]
        _lib.orc_so3_log.argtypes = [dp, dp]
        _lib.orc_left_jacobian.argtypes = [dp, dp]
        _lib.orc_calc_body_cov.argtypes = [dp, C.c_double, C.c_double, dp]
        _lib.orc_boxplus.argtypes = [C.POINTER(VmpState), dp]
        _lib.orc_boxminus.argtypes = [C.POINTER(VmpState), C.POINTER(VmpState), dp]
        _lib.orc_predict.argtypes = [C.c_void_p, dp, dp, C.c_double]
        _lib.orc_voxel_key.argtypes = [C.c_void_p, dp, C.POINTER(C.c_int64)]
        _lib.orc_set_threads.argtypes = [C.c_void_p, C.c_int]
        _lib.orc_lio_process.argtypes = [C.c_void_p, C.POINTER(VmpImu), C.c_int, C.POINTER(C.c_float), C.c_int,
                                         C.c_double, C.c_double, C.POINTER(VmpScanStats)]
        _lib.orc_lio_state.argtypes = [C.c_void_p, C.POINTER(VmpState), dp, C.POINTER(C.c_int)]
        _lib.orc_map_update_timed.argtypes = [C.c_void_p, dp, dp, C.c_int, C.POINTER(VmpUpdateStats)]
        _lib.orc_map_update_timed.restype = C.c_double
        _lib.orc_get_prior.argtypes = [C.c_void_p, C.POINTER(VmpState), dp]
        _lib.orc_get_iter_state.argtypes = [C.c_void_p, C.c_int, C.POINTER(VmpState)]
    return _lib


class Oracle(HotPath):
    """The CPU oracle behind the same Python surface as the CUDA path (prefix orc_)."""

    def __init__(self, cfg):
        super().__init__(cfg, lib=lib(), prefix="orc_")

    def set_threads(self, n: int):
        self._lib.orc_set_threads(self._h, n)

    def lio_process(self, imus: np.ndarray, cloud_xyzc: np.ndarray, t0: float, t1: float):
        """LIOBuilder::process(SyncPackage&) (lio_builder.cpp:175-248); cloud is sorted/undistorted in place."""
        assert cloud_xyzc.dtype == np.float32 and cloud_xyzc.flags.c_contiguous
        st = VmpScanStats()
        self._n = cloud_xyzc.shape[0]
        self._check(self._lib.orc_lio_process(self._h, imus.ctypes.data_as(C.POINTER(VmpImu)), imus.shape[0],
                                              fptr(cloud_xyzc), cloud_xyzc.shape[0], t0, t1, C.byref(st)))
        if self.cfg.scan_resolution > 0 and st.iters > 0:
            self._n = self.lidar_cloud().shape[0]
        return st

    def track_margins(self, on: bool = True):
        """Switch the gate-margin instrumentation on (off by default: the timed CPU baselines run without it)."""
        self._lib.orc_track_margins.argtypes = [C.c_void_p, C.c_int]
        self._lib.orc_track_margins(self._h, 1 if on else 0)

    def gate_margins(self) -> dict:
        """Smallest |margin| to the threshold over every gate decision so far (plane fit, 3-sigma gate, merge angle / distance)."""
        out = np.zeros(4)
        self._lib.orc_gate_margins.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        self._lib.orc_gate_margins(self._h, dptr(out))
        return dict(zip(("plane", "gate", "merge_angle", "merge_dist"), out.tolist()))

    def lio_state(self):
        x = VmpState()
        P = np.zeros((23, 23))
        s = C.c_int(0)
        self._lib.orc_lio_state(self._h, C.byref(x), dptr(P), C.byref(s))
        return x, P, s.value

    def get_prior(self):
        x = VmpState()
        P = np.zeros((23, 23))
        self._lib.orc_get_prior(self._h, C.byref(x), dptr(P))
        return x, P

    def get_iter_state(self, k: int):
        x = VmpState()
        if self._lib.orc_get_iter_state(self._h, k, C.byref(x)) != 0:
            raise IndexError(k)
        return x

    def get_iter_Hb(self, k: int):
        H = np.zeros((12, 12))
        b = np.zeros(12)
        self._lib.orc_get_iter_Hb.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        if self._lib.orc_get_iter_Hb(self._h, k, dptr(H), dptr(b)) != 0:
            raise IndexError(k)
        return H, b

    def predict(self, acc, gyro, dt):
        a = np.ascontiguousarray(acc, np.float64)
        g = np.ascontiguousarray(gyro, np.float64)
        self._lib.orc_predict(self._h, dptr(a), dptr(g), dt)

    def voxel_key(self, p):
        p = np.ascontiguousarray(p, np.float64)
        k = np.zeros(3, np.int64)
        self._lib.orc_voxel_key(self._h, dptr(p), k.ctypes.data_as(C.POINTER(C.c_int64)))
        return k

    def map_update_timed(self, pts_world, cov):
        p = np.ascontiguousarray(pts_world, np.float64).reshape(-1, 3)
        c = np.ascontiguousarray(cov, np.float64).reshape(-1, 9)
        st = VmpUpdateStats()
        sec = self._lib.orc_map_update_timed(self._h, dptr(p), dptr(c), p.shape[0], C.byref(st))
        return sec, st.as_dict()


def std_build_voxels(cloud_xyzi, voxel_size=1.0, voxel_min_point=10, voxel_plane_thresh=0.01):
    """orc_std_build_voxels: STDManager::buildVoxels restated (descriptor.cpp:70-122)."""
    from voxelmapplus_fastlio2_b200.bindings import std_build_voxels as _call
    return _call(cloud_xyzi, voxel_size, voxel_min_point, voxel_plane_thresh, lib=lib(), prefix="orc_")


# ---- small math, for unit tests against numpy -------------------------------------
def eig3(A):
    A = np.ascontiguousarray(A, np.float64).reshape(3, 3)
    w = np.zeros(3)
    V = np.zeros((3, 3))
    lib().orc_eig3(dptr(A), dptr(w), dptr(V))
    return w, V


def inverse23(A):
    A = np.ascontiguousarray(A, np.float64).reshape(23, 23)
    out = np.zeros((23, 23))
    lib().orc_inverse23(dptr(A), dptr(out))
    return out


def so3_exp(w):
    w = np.ascontiguousarray(w, np.float64)
    R = np.zeros((3, 3))
    lib().orc_so3_exp(dptr(w), dptr(R))
    return R


def so3_log(R):
    R = np.ascontiguousarray(R, np.float64).reshape(3, 3)
    w = np.zeros(3)
    lib().orc_so3_log(dptr(R), dptr(w))
    return w


def left_jacobian(w):
    w = np.ascontiguousarray(w, np.float64)
    J = np.zeros((3, 3))
    lib().orc_left_jacobian(dptr(w), dptr(J))
    return J


def calc_body_cov(p, range_inc=0.04, degree_inc=0.1):
    p = np.array(p, np.float64)
    cov = np.zeros((3, 3))
    lib().orc_calc_body_cov(dptr(p), range_inc, degree_inc, dptr(cov))
    return p, cov


def boxplus(x: VmpState, delta):
    x = x.copy()
    d = np.ascontiguousarray(delta, np.float64)
    lib().orc_boxplus(C.byref(x), dptr(d))
    return x


def boxminus(a: VmpState, b: VmpState):
    d = np.zeros(23)
    lib().orc_boxminus(C.byref(a), C.byref(b), dptr(d))
    return d
