"""GPU parity at the sizes BASELINE.json quotes (C1 .. C4), through the C ABI against the CPU oracle.  The procedures live in
tests/parity_cases.py; tools/parity_run.py runs the same ones at the full lengths (1000 / 200 / 10 000 scans, 50 M points) and
writes the JSON summary committed under profiles/.  Sizes here are chosen so that the five tests together stay within a few
minutes of a B200 box (the oracle on the host is the slow side); VMP_PARITY_SCALE=<float> scales the lengths."""
import os

import pytest

import parity_cases as pc

pytestmark = pytest.mark.gpu
SCALE = float(os.environ.get("VMP_PARITY_SCALE", "1"))


def test_c1_free_running_300_scans(oracle_mod):
    """configs[0] as specified: 20 000 pts, 0.5 m voxels, >= 300 scans free-running: tier 3 (1 mm / 0.01 deg), iterations equal."""
    r = pc.c1_free_running(oracle_mod, scans=int(300 * SCALE))
    print(r)


def test_c1_teacher_forced_300_scans(oracle_mod):
    """configs[0] size, teacher-forced: keys / status bit-exact in every pass, H / b <= 1e-9, map bit-exact."""
    r = pc.c1_teacher_forced(oracle_mod, scans=int(300 * SCALE))
    print(r)


def test_c2_teacher_forced_50_scans(oracle_mod):
    """configs[1]: 200 000 pts, 0.25 m voxels, 4 iterations, >= 50 scans teacher-forced."""
    r = pc.c2_teacher_forced(oracle_mod, scans=int(50 * SCALE))
    print(r)


def test_c3_city_2000_scans_capacity_100k(oracle_mod):
    """configs[2]: city drive at map_capacity 100 000, >= 2000 scans: evicted keys per scan and final map bit-exact (the map is
    full around scan 600; ~450 000 evictions afterwards).  tools/parity_run.py runs the same procedure over the full 10 000 scans."""
    r = pc.c3_city_eviction(oracle_mod, scans=int(2000 * SCALE))
    assert SCALE < 1 or r["evicted"] > 100000, r
    print(r)


def test_c4_map_slice_5m_points(oracle_mod):
    """configs[3] slice: >= 5 M points through VoxelMap::update on both sides, counters + evicted keys + final map bit-exact."""
    r = pc.c4_map_slice(oracle_mod, points=int(5_000_000 * SCALE))
    print(r)
