"""Build the CUDA library in-tree: voxelmapplus_fastlio2_b200/libvmp_b200.so (sm_100a only).

nvcc cross-compiles without a GPU.  Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only, no PTX for other targets
  -fmad=false                               no FMA contraction on the device: voxel keys, plane fits
                                            and gate decisions must round exactly like the CPU oracle
  -Xcompiler -ffp-contract=off              same contract for the host-side estimator code
  -lineinfo                                 so ncu's source page maps to these files
"""
from __future__ import annotations

import os
import subprocess
import sys

_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_DIR, "csrc")
LIB = os.path.join(_DIR, "libvmp_b200.so")
SOURCES = ["vmp_iekf.cu", "vmp_map.cu", "vmp_downsample.cu", "vmp_readback.cu", "vmp_std.cu", "vmp_capi.cu", "vmp_lio.cu"]
HEADERS = ["vmp_math.cuh", "vmp_state.cuh", "vmp_device.cuh", "vmp_kernels.h", "vmp_lio.hpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def nvcc_cmd(extra=()):
    return [NVCC, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
            "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function,-Wno-unknown-pragmas",
            "-Xptxas", "-v" if os.environ.get("VMP_PTXAS_V") else "-warn-spills",
            *extra, "-shared", "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_DIR, "..", "include", "vmp_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = nvcc_cmd()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libvmp_b200.so")
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
