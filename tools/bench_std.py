"""Wall time of vmp_std_build_voxels against the oracle restatement on the same cloud.
usage: python tools/bench_std.py [n_points]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import oracle_py
from test_std_voxels import submap_cloud
import ctypes

from voxelmapplus_fastlio2_b200.bindings import load_library, std_build_voxels

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cloud = submap_cloud(seed=7, n=n, extent=120.0)
std_build_voxels(cloud[:1000])                      # context + module load
lib = load_library()
lib.vmp_std_last_device_ms.restype = ctypes.c_double
t, d = [], []
for _ in range(5):
    t0 = time.perf_counter(); v = std_build_voxels(cloud); t.append(time.perf_counter() - t0); d.append(lib.vmp_std_last_device_ms())
t0 = time.perf_counter(); o = oracle_py.std_build_voxels(cloud); t_cpu = time.perf_counter() - t0
print(json.dumps({"points": n, "voxels": len(v), "planes": int((v["flags"] & 2).astype(bool).sum()),
                  "device_kernels_ms_min": round(min(d), 3), "device_call_ms_min": round(min(t) * 1e3, 3), "device_call_ms_median": round(sorted(t)[2] * 1e3, 3),
                  "oracle_1thread_ms": round(t_cpu * 1e3, 1), "same_voxels": bool(np.array_equal(v["key"], o["key"]))}))
