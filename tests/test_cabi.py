"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/vmp_b200.h
declares, mirrors the header's POD layouts, and fails LOUDLY (no CPU fallback) when there is no B200."""
import ctypes as C
import os
import re

import pytest

from voxelmapplus_fastlio2_b200 import bindings
from voxelmapplus_fastlio2_b200.ctypes_defs import (K_COUNT, VmpConfig, VmpPlane, VmpScanStats, VmpState, VmpUpdateStats,
                                                    default_config)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vmp_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vmp_[a-z0-9_]+)\s*\(", src)))


def test_header_cites_reference_interfaces():
    src = open(HEADER).read()
    for cite in ("voxel_map.cpp:200-230", "voxel_map.cpp:232-256", "lio_builder.cpp:250-311", "ieskf.cpp:125-156",
                 "lio_builder.h:17-42", "ieskf.h:31-62", "voxel_map.h:43-51", "commons.cpp:18-45"):
        assert cite in src, f"header must cite {cite}"


def test_library_exports_every_declared_symbol():
    lib = bindings.load_library()
    names = declared_functions()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"libvmp_b200.so does not export {missing}"


def test_library_does_not_link_the_oracle():
    """the product must not depend on oracle/ in any way"""
    import subprocess
    out = subprocess.run(["ldd", bindings.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", bindings.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in syms
    for root, _, files in os.walk(os.path.join(ROOT, "voxelmapplus_fastlio2_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "oracle_py" not in txt and "liboracle" not in txt and '"../oracle' not in txt and "from oracle" not in txt, f


def test_struct_layouts_match_header():
    assert C.sizeof(VmpState) == 36 * 8
    assert C.sizeof(VmpUpdateStats) == 13 * 8
    assert C.sizeof(VmpPlane) == 3 * 8 + (3 + 9 + 3 + 36 + 3) * 8 + 4 * 4 + 2 * 8
    assert C.sizeof(VmpScanStats) == 4 + 32 + 4 + 104 + 4 + 4
    from voxelmapplus_fastlio2_b200.ctypes_defs import VmpStdVoxel
    assert C.sizeof(VmpStdVoxel) == 3 * 8 + 4 + 4 + (3 + 9 + 3 + 3 + 9) * 8      # static_assert of the same number in vmp_std.cu
    lib = bindings.load_library()
    lib.vmp_config_default.argtypes = [C.POINTER(VmpConfig)]
    c = VmpConfig()
    lib.vmp_config_default(C.byref(c))
    d = default_config()
    for name, _ in VmpConfig._fields_:
        a, b = getattr(c, name), getattr(d, name)
        if name in ("scan_resolution",):        # header default = reference default 0.1; the python helper turns it off
            assert a == pytest.approx(0.1) and b == 0.0
            continue
        if hasattr(a, "__len__"):
            assert list(a) == list(b), name
        else:
            assert a == b, name
    lib.vmp_kernel_name.restype = C.c_char_p
    lib.vmp_kernel_name.argtypes = [C.c_int]
    names = [lib.vmp_kernel_name(k).decode() for k in range(K_COUNT)]
    assert names[0] == "k_scan_in" and names[-1] == "k_downsample" and names[-2] == "k_undistort" and "?" not in names
    assert lib.vmp_kernel_name(K_COUNT).decode() == "?"


def test_no_cpu_fallback_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(bindings.VmpError, match="no CUDA device|CPU fallback"):
        bindings.HotPath(default_config(max_points_per_scan=64, map_capacity=64))


def test_invalid_config_is_rejected_before_touching_the_device():
    lib = bindings.load_library()
    lib.vmp_create.argtypes = [C.POINTER(VmpConfig), C.POINTER(C.c_void_p)]
    lib.vmp_last_error.restype = C.c_char_p
    h = C.c_void_p()
    bad = default_config(voxel_size=-1.0)
    assert lib.vmp_create(C.byref(bad), C.byref(h)) == -1
    assert b"invalid configuration" in lib.vmp_last_error()
    assert lib.vmp_create(None, C.byref(h)) == -1
