"""Host logic of the scan upload (csrc/vmp_stage.hpp): multi-threaded staging copy + time-order check, on the CPU."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stage_pool_copies_and_order_check(tmp_path):
    exe = str(tmp_path / "stage_pool_check")
    src = os.path.join(ROOT, "tests", "cpp", "stage_pool_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, src], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-500:]
