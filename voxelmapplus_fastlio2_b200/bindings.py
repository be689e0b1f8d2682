"""ctypes binding of the C ABI in include/vmp_b200.h.

`HotPath` is a thin, numpy-in / numpy-out wrapper over one `vmp_handle` (one trajectory:
one voxel map + one filter, resident on one B200).  The method names follow the
reference's own surface (`VoxelMap::build/update`, `LIOBuilder::sharedUpdateFunc`,
`IESKF::update` + `map->update` as `scan`), see the header for file:line citations.

The same class binds any library that exports the same entry points under another
prefix; tests use that to drive the CPU oracle (prefix ``orc_``) with identical calls.
There is no fallback of any kind in here: if the CUDA library is missing or there is no
B200, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .ctypes_defs import (IMU_DTYPE, PLANE_DTYPE, STD_VOXEL_DTYPE, VmpConfig, VmpImu, VmpPlane, VmpScanStats, VmpState,
                          VmpStdVoxel, VmpUpdateStats, dptr, fptr)

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libvmp_b200.so")

_libs: dict = {}


class VmpError(RuntimeError):
    pass


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """dlopen the CUDA library. Fails loudly when it has not been built (no CPU fallback)."""
    if path not in _libs:
        if not os.path.exists(path):
            raise VmpError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        _libs[path] = C.CDLL(path, mode=C.RTLD_GLOBAL)
    return _libs[path]


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class HotPath:
    """One trajectory's map + filter behind the C ABI."""

    def __init__(self, cfg: VmpConfig, lib: C.CDLL | None = None, prefix: str = "vmp_"):
        self._lib = lib if lib is not None else load_library()
        self._p = prefix
        self.cfg = cfg
        self._h = C.c_void_p()
        self._declare()
        self._check(self._fn("create")(C.byref(cfg), C.byref(self._h)))

    # -- plumbing --------------------------------------------------------------
    def _fn(self, name):
        return getattr(self._lib, self._p + name)

    def _declare(self):
        vp, ip, dp, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_float)
        sig = {
            "create": [C.POINTER(VmpConfig), C.POINTER(C.c_void_p)],
            "destroy": [vp],
            "map_build": [vp, dp, dp, C.c_int, C.POINTER(VmpUpdateStats)],
            "map_update": [vp, dp, dp, C.c_int, C.POINTER(VmpUpdateStats)],
            "measure": [vp, C.POINTER(VmpState), dp, dp, dp, ip],
            "set_scan": [vp, fp, C.c_int],
            "scan": [vp, C.POINTER(VmpState), dp, fp, C.c_int, C.POINTER(VmpScanStats)],
            "set_state": [vp, C.POINTER(VmpState), dp],
            "get_state": [vp, C.POINTER(VmpState), dp],
            "first_scan": [vp, C.POINTER(VmpState), dp, fp, C.c_int, C.POINTER(VmpUpdateStats)],
            "dump_correspondences": [vp, C.POINTER(C.c_int64), C.POINTER(C.c_uint8), dp, dp, C.c_int],
            "dump_world_points": [vp, dp, dp, C.c_int],
            "dump_map": [vp, C.POINTER(VmpPlane), C.c_int, ip],
            "dump_evicted": [vp, C.POINTER(C.c_int64), C.c_int, ip],
            "map_size": [vp, ip],
            "downsample": [vp, fp, C.c_int, C.c_double, fp, C.c_int, ip],
            "get_lidar_cloud": [vp, fp, C.c_int, ip],
        }
        for name, args in sig.items():
            f = self._fn(name)
            f.argtypes = args
            f.restype = C.c_int
        if self._p == "vmp_":
            self._lib.vmp_last_error.restype = C.c_char_p
            self._lib.vmp_scan_dev.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(VmpScanStats)]
            self._lib.vmp_scan_dev.restype = C.c_int
            self._lib.vmp_launch_count.argtypes = [vp]
            self._lib.vmp_launch_count.restype = C.c_int64

    def _check(self, rc):
        if rc != 0:
            msg = ""
            if self._p == "vmp_":
                msg = (self._lib.vmp_last_error() or b"").decode()
            raise VmpError(f"{self._p}* call failed with status {rc}: {msg}")

    def close(self):
        if self._h:
            self._fn("destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- S2: VoxelMap ----------------------------------------------------------
    def map_build(self, pts_world, cov) -> dict:
        """VoxelMap::build (voxel_map.cpp:200-230)."""
        p, c = _f64(pts_world, (-1, 3)), _f64(cov, (-1, 9))
        st = VmpUpdateStats()
        self._check(self._fn("map_build")(self._h, dptr(p), dptr(c), p.shape[0], C.byref(st)))
        return st.as_dict()

    def map_update(self, pts_world, cov) -> dict:
        """VoxelMap::update (voxel_map.cpp:232-256)."""
        p, c = _f64(pts_world, (-1, 3)), _f64(cov, (-1, 9))
        st = VmpUpdateStats()
        self._check(self._fn("map_update")(self._h, dptr(p), dptr(c), p.shape[0], C.byref(st)))
        return st.as_dict()

    # -- S1: measurement plug-in -------------------------------------------------
    def set_scan(self, pts_lidar):
        """lio_builder.cpp:224-229."""
        p = np.ascontiguousarray(pts_lidar, dtype=np.float32).reshape(-1, 3)
        self._n = p.shape[0]
        self._check(self._fn("set_scan")(self._h, fptr(p), p.shape[0]))

    def measure(self, x: VmpState, P):
        """One LIOBuilder::sharedUpdateFunc call (lio_builder.cpp:250-311) -> (H 12x12, b 12, effect_num)."""
        P = _f64(P, (23, 23))
        H = np.zeros((12, 12))
        b = np.zeros(12)
        e = C.c_int(0)
        self._check(self._fn("measure")(self._h, C.byref(x), dptr(P), dptr(H), dptr(b), C.byref(e)))
        return H, b, e.value

    # -- S3: whole timed region --------------------------------------------------
    def scan(self, x: VmpState, P, pts_lidar):
        """lio_builder.cpp:224-246 -> (posterior x, posterior P, VmpScanStats)."""
        p = np.ascontiguousarray(pts_lidar, dtype=np.float32).reshape(-1, 3)
        self._n = p.shape[0]
        x = x.copy()
        P = _f64(P, (23, 23)).copy()
        st = VmpScanStats()
        self._check(self._fn("scan")(self._h, C.byref(x), dptr(P), fptr(p), p.shape[0], C.byref(st)))
        return x, P, st

    def scan_staged(self, x: VmpState, P, pts_lidar):
        """vmp_scan_buffer + vmp_scan_staged: the points are written straight into the handle's pinned staging area."""
        p = np.asarray(pts_lidar, dtype=np.float32).reshape(-1, 3)
        n = p.shape[0]
        self._lib.vmp_scan_buffer.argtypes = [C.c_void_p]
        self._lib.vmp_scan_buffer.restype = C.POINTER(C.c_float)
        self._lib.vmp_scan_staged.argtypes = [C.c_void_p, C.POINTER(VmpState), C.POINTER(C.c_double), C.c_int, C.POINTER(VmpScanStats)]
        self._lib.vmp_scan_staged.restype = C.c_int
        buf = self._lib.vmp_scan_buffer(self._h)
        np.ctypeslib.as_array(buf, shape=(max(n, 1), 3))[:n] = p
        self._n = n
        x = x.copy()
        P = _f64(P, (23, 23)).copy()
        st = VmpScanStats()
        self._check(self._lib.vmp_scan_staged(self._h, C.byref(x), dptr(P), n, C.byref(st)))
        return x, P, st

    def scan_filled(self, x: VmpState, P, records, stride=None):
        """vmp_scan_buffer_fill + vmp_scan_staged: x y z are taken out of the caller's point records (n x stride float32) by the handle's staging helpers."""
        r = np.ascontiguousarray(records, dtype=np.float32)
        stride = r.shape[1] if stride is None else stride
        n = r.shape[0]
        self._lib.vmp_scan_buffer_fill.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int]
        self._lib.vmp_scan_buffer_fill.restype = C.c_int
        self._lib.vmp_scan_staged.argtypes = [C.c_void_p, C.POINTER(VmpState), C.POINTER(C.c_double), C.c_int, C.POINTER(VmpScanStats)]
        self._lib.vmp_scan_staged.restype = C.c_int
        self._check(self._lib.vmp_scan_buffer_fill(self._h, fptr(r), stride, n))
        self._n = n
        x = x.copy()
        P = _f64(P, (23, 23)).copy()
        st = VmpScanStats()
        self._check(self._lib.vmp_scan_staged(self._h, C.byref(x), dptr(P), n, C.byref(st)))
        return x, P, st

    def first_scan(self, x: VmpState, P, pts_lidar) -> dict:
        """MAP_INIT branch (lio_builder.cpp:185-211)."""
        p = np.ascontiguousarray(pts_lidar, dtype=np.float32).reshape(-1, 3)
        self._n = p.shape[0]
        P = _f64(P, (23, 23))
        st = VmpUpdateStats()
        self._check(self._fn("first_scan")(self._h, C.byref(x), dptr(P), fptr(p), p.shape[0], C.byref(st)))
        return st.as_dict()

    def get_prior(self):
        """prior (x, P) of the last scan as the device saw it (uploaded, or propagated on the device by vmp_scan_raw_predict)"""
        x = VmpState()
        P = np.zeros((23, 23))
        self._lib.vmp_get_prior.argtypes = [C.c_void_p, C.POINTER(VmpState), C.POINTER(C.c_double)]
        self._check(self._lib.vmp_get_prior(self._h, C.byref(x), dptr(P)))
        return x, P

    def set_state(self, x: VmpState, P):
        P = _f64(P, (23, 23))
        self._check(self._fn("set_state")(self._h, C.byref(x), dptr(P)))

    def get_state(self):
        x = VmpState()
        P = np.zeros((23, 23))
        self._check(self._fn("get_state")(self._h, C.byref(x), dptr(P)))
        return x, P

    # -- pre-processing ----------------------------------------
    def downsample(self, cloud_xyzc, leaf: float) -> np.ndarray:
        """scan_filter.filter(): pcl::VoxelGrid leaf-centroid filter (lio_builder.cpp:13-14, 215-219) -> M x 4 float32."""
        c = np.ascontiguousarray(cloud_xyzc, dtype=np.float32).reshape(-1, 4)
        out = np.zeros((max(c.shape[0], 1), 4), np.float32)
        m = C.c_int(0)
        self._check(self._fn("downsample")(self._h, fptr(c), c.shape[0], float(leaf), fptr(out), c.shape[0], C.byref(m)))
        return out[:m.value].copy()

    def lidar_cloud(self, cap: int | None = None) -> np.ndarray:
        """LIOBuilder::lidar_cloud of the last raw scan (the filter input; leaf centroids when scan_resolution > 0)."""
        cap = self.cfg.max_points_per_scan if cap is None else cap
        out = np.zeros((max(cap, 1), 4), np.float32)
        m = C.c_int(0)
        self._check(self._fn("get_lidar_cloud")(self._h, fptr(out), cap, C.byref(m)))
        return out[:m.value].copy()

    # -- observation -------------------------------------------------------------
    def dump_correspondences(self, n=None):
        n = self._n if n is None else n
        keys = np.zeros((n, 3), np.int64)
        status = np.zeros(n, np.uint8)
        res = np.zeros(n)
        norm = np.zeros((n, 3))
        self._check(self._fn("dump_correspondences")(
            self._h, keys.ctypes.data_as(C.POINTER(C.c_int64)), status.ctypes.data_as(C.POINTER(C.c_uint8)),
            dptr(res), dptr(norm), n))
        return {"keys": keys, "status": status, "residual": res, "plane_norm": norm}

    def dump_world_points(self, n=None):
        n = self._n if n is None else n
        pts = np.zeros((n, 3))
        cov = np.zeros((n, 9))
        self._check(self._fn("dump_world_points")(self._h, dptr(pts), dptr(cov), n))
        return pts, cov

    def map_size(self) -> int:
        c = C.c_int(0)
        self._check(self._fn("map_size")(self._h, C.byref(c)))
        return c.value

    def dump_map(self) -> np.ndarray:
        """All live voxels in LRU order (front first) as a structured array (PLANE_DTYPE)."""
        n = self.map_size()
        out = np.zeros(max(n, 1), PLANE_DTYPE)
        c = C.c_int(0)
        self._check(self._fn("dump_map")(self._h, out.ctypes.data_as(C.POINTER(VmpPlane)), n, C.byref(c)))
        return out[:min(n, c.value)]

    def dump_evicted(self, cap=1 << 20) -> np.ndarray:
        keys = np.zeros((cap, 3), np.int64)
        c = C.c_int(0)
        self._check(self._fn("dump_evicted")(self._h, keys.ctypes.data_as(C.POINTER(C.c_int64)), cap, C.byref(c)))
        return keys[:c.value].copy()

    # -- product-only ------------------------------------------------------------
    def scan_dev(self, pts_dev_ptr: int, n: int, prior_dev_ptr: int | None = None) -> VmpScanStats:
        """vmp_scan_dev: scan (n x 3 float32) and optionally the prior (36+529 float64) already resident in HBM."""
        st = VmpScanStats()
        self._n = n
        self._check(self._lib.vmp_scan_dev(self._h, C.c_void_p(pts_dev_ptr),
                                           C.c_void_p(prior_dev_ptr) if prior_dev_ptr else None, n, C.byref(st)))
        return st

    def set_pipelined(self, on: bool = True):
        """vmp_set_pipelined: scan() returns with the posterior while the map update of that scan is still running."""
        self._lib.vmp_set_pipelined.argtypes = [C.c_void_p, C.c_int]
        self._check(self._lib.vmp_set_pipelined(self._h, 1 if on else 0))

    def sync(self):
        self._lib.vmp_sync.argtypes = [C.c_void_p]
        self._check(self._lib.vmp_sync(self._h))

    def launch_count(self) -> int:
        return int(self._lib.vmp_launch_count(self._h))

    def debug_counters(self):
        out = (C.c_int * 8)()
        self._lib.vmp_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        self._check(self._lib.vmp_debug_counters(self._h, out))
        return list(out)

    def profile_enable(self, on: bool = True):
        """Per-kernel CUDA-event timing: scans run the same kernels one by one instead of the graph."""
        self._lib.vmp_profile_enable.argtypes = [C.c_void_p, C.c_int]
        self._check(self._lib.vmp_profile_enable(self._h, 1 if on else 0))

    def profile_reset(self):
        self._lib.vmp_profile_reset.argtypes = [C.c_void_p]
        self._check(self._lib.vmp_profile_reset(self._h))

    def profile_read(self) -> dict:
        """{kernel name: (accumulated ms, launches)} since the last reset."""
        from .ctypes_defs import K_COUNT
        ms = np.zeros(K_COUNT)
        cnt = np.zeros(K_COUNT, np.int64)
        self._lib.vmp_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        self._lib.vmp_kernel_name.argtypes = [C.c_int]
        self._lib.vmp_kernel_name.restype = C.c_char_p
        self._check(self._lib.vmp_profile_read(self._h, dptr(ms), cnt.ctypes.data_as(C.POINTER(C.c_int64))))
        return {self._lib.vmp_kernel_name(k).decode(): (float(ms[k]), int(cnt[k])) for k in range(K_COUNT)}


def std_build_voxels(cloud_xyzi, voxel_size: float = 1.0, voxel_min_point: int = 10, voxel_plane_thresh: float = 0.01,
                     lib: C.CDLL | None = None, prefix: str = "vmp_") -> np.ndarray:
    """STDManager::buildVoxels (std_matcher/src/std_manager/descriptor.cpp:70-122; defaults: descriptor.h Config) on the device.
    Returns one STD_VOXEL_DTYPE record per voxel, in the order of the voxels' first points."""
    lib = lib or load_library()
    fn = getattr(lib, prefix + "std_build_voxels")
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_double, C.c_int, C.c_double, C.POINTER(VmpStdVoxel), C.c_int, C.POINTER(C.c_int)]
    cloud = np.ascontiguousarray(cloud_xyzi, np.float32).reshape(-1, 4)
    n = cloud.shape[0]
    out = np.zeros(max(n, 1), STD_VOXEL_DTYPE)
    c = C.c_int(0)
    rc = fn(cloud.ctypes.data_as(C.POINTER(C.c_float)), n, voxel_size, voxel_min_point, voxel_plane_thresh,
            out.ctypes.data_as(C.POINTER(VmpStdVoxel)), n, C.byref(c))
    if rc != 0:
        msg = ""
        if prefix == "vmp_":
            lib.vmp_last_error.restype = C.c_char_p
            msg = (lib.vmp_last_error() or b"").decode()
        raise VmpError(f"{prefix}std_build_voxels: rc={rc} {msg}")
    return out[:c.value].copy()


def imu_array(n: int) -> np.ndarray:
    return np.zeros(n, IMU_DTYPE)
