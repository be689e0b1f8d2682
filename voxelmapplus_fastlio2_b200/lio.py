"""Python view of the host-side C++ estimator `vmp::LIOBuilder` (csrc/vmp_lio.hpp), the mirror of the
reference's `lio::LIOBuilder` (lio_builder.h:56-83): `loadConfig` happens in the constructor,
`process(SyncPackage)` per scan, `kf.x()` / `kf.P()` / `status` through `state()`."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .bindings import HotPath, VmpError, load_library
from .ctypes_defs import VmpConfig, VmpImu, VmpScanStats, VmpState, dptr, fptr

IMU_INIT, MAP_INIT, LIO_MAPPING = 0, 1, 2


class _BorrowedHotPath(HotPath):
    """HotPath view over the vmp_handle owned by a LIOBuilder (never destroys it)."""

    def __init__(self, lib, cfg, handle):
        self._lib = lib
        self._p = "vmp_"
        self.cfg = cfg
        self._h = C.c_void_p(handle)
        self._declare()

    def close(self):
        self._h = C.c_void_p()


class LIOBuilder:
    def __init__(self, cfg: VmpConfig, pipelined: bool = False, device_undistort: bool = True, cloud_writeback: bool = True,
                 device_predict: bool = False):
        """pipelined: process() returns with the posterior of the scan while its map update still runs on the device
        (vmp_set_pipelined); the map counters in the returned stats then describe the previous scan."""
        self._lib = load_library()
        L = self._lib
        L.vmp_lio_create.argtypes = [C.POINTER(VmpConfig), C.POINTER(C.c_void_p)]
        L.vmp_lio_destroy.argtypes = [C.c_void_p]
        L.vmp_lio_process.argtypes = [C.c_void_p, C.POINTER(VmpImu), C.c_int, C.POINTER(C.c_float), C.c_int,
                                      C.c_double, C.c_double, C.POINTER(VmpScanStats)]
        L.vmp_lio_state.argtypes = [C.c_void_p, C.POINTER(VmpState), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.vmp_lio_map.argtypes = [C.c_void_p]
        L.vmp_lio_map.restype = C.c_void_p
        L.vmp_last_error.restype = C.c_char_p
        self.cfg = cfg
        self._h = C.c_void_p()
        self._check(L.vmp_lio_create(C.byref(cfg), C.byref(self._h)))
        self.map = _BorrowedHotPath(L, cfg, L.vmp_lio_map(self._h))
        if pipelined:
            self.map.set_pipelined(True)
        if device_predict:              # IESKF::predict on the device as well (vmp_scan_raw_predict): state / P resident there
            L.vmp_lio_set_device_predict.argtypes = [C.c_void_p, C.c_int]
            self._check(L.vmp_lio_set_device_predict(self._h, 1))
        if not cloud_writeback:         # process() leaves the caller's cloud as it is; the compensated cloud stays readable (lidar_cloud)
            L.vmp_lio_set_cloud_writeback.argtypes = [C.c_void_p, C.c_int]
            self._check(L.vmp_lio_set_cloud_writeback(self._h, 0))
        if not device_undistort:        # default: lio_builder.cpp:127-152 runs on the device, in the scan's graph
            L.vmp_lio_set_device_undistort.argtypes = [C.c_void_p, C.c_int]
            self._check(L.vmp_lio_set_device_undistort(self._h, 0))

    def _check(self, rc):
        if rc != 0:
            raise VmpError(f"vmp_lio_* failed with status {rc}: {(self._lib.vmp_last_error() or b'').decode()}")

    def process(self, imus: np.ndarray, cloud_xyzc: np.ndarray, t0: float, t1: float) -> VmpScanStats:
        """LIOBuilder::process (lio_builder.cpp:175-248). The cloud is sorted + undistorted in place."""
        assert cloud_xyzc.dtype == np.float32 and cloud_xyzc.flags.c_contiguous
        st = VmpScanStats()
        self.map._n = cloud_xyzc.shape[0]
        self._check(self._lib.vmp_lio_process(self._h, imus.ctypes.data_as(C.POINTER(VmpImu)), imus.shape[0],
                                              fptr(cloud_xyzc), cloud_xyzc.shape[0], t0, t1, C.byref(st)))
        if self.cfg.scan_resolution > 0 and st.iters > 0 and st.map.n_points > 0:
            self.map._n = int(st.map.n_points)      # the filter saw the leaf centroids (synchronous mode)
        return st

    def state(self):
        x = VmpState()
        P = np.zeros((23, 23))
        s = C.c_int(0)
        self._check(self._lib.vmp_lio_state(self._h, C.byref(x), dptr(P), C.byref(s)))
        return x, P, s.value

    def prior(self):
        """The (x, P) that the last process() handed to the device update (after IMU propagation)."""
        x = VmpState()
        P = np.zeros((23, 23))
        self._lib.vmp_lio_prior.argtypes = [C.c_void_p, C.POINTER(VmpState), C.POINTER(C.c_double)]
        self._check(self._lib.vmp_lio_prior(self._h, C.byref(x), dptr(P)))
        return x, P

    def close(self):
        if self._h:
            self.map.close()
            self._lib.vmp_lio_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
