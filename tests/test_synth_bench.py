"""Host-side pieces around the measurement: the synthetic driver (culling, parallel generation) and bench.py's accounting."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from voxelmapplus_fastlio2_b200 import synth  # noqa: E402


def test_culled_cast_equals_full_cast():
    """Scene.near() only drops rectangles no ray of the scan can reach: the culled cast returns the same ranges."""
    sc = synth.scene_city(pilasters=True)
    traj = synth.Trajectory(centre=(600.0, 600.0, 1.8), ax=560.0, ay=2.0, period=448.0)
    full = synth.Sequence(scene=sc, traj=traj, sensor=synth.SensorConfig(pts_per_scan=1500), seed=7)
    cull = synth.Sequence(scene=sc, traj=traj, sensor=synth.SensorConfig(pts_per_scan=1500), seed=7, cull=True)
    for idx in (0, 400, 2240, 5000):
        a, _ = full.cloud(idx)
        b, _ = cull.cloud(idx)
        assert a.shape == b.shape and a.shape[0] > 500
        assert np.allclose(a, b, rtol=0, atol=1e-5)          # same hits (BLAS may round the last bit differently per shape)
    near = sc.near(np.array([[600.0, 600.0, 1.8]]), 40.0)
    assert 0 < near.c.shape[0] < sc.c.shape[0] // 4


def test_city_pilasters_keep_the_base_scene():
    a, b = synth.scene_city(), synth.scene_city(pilasters=True)
    n = a.c.shape[0]
    assert b.c.shape[0] > n and np.array_equal(a.c, b.c[:n]) and np.array_equal(a.n, b.n[:n])


def test_parallel_generation_is_deterministic():
    import bench
    wl = dict(bench.WORKLOADS["c1"], pts=800)
    par = bench.make_packages(wl, 11, 5, workers=2)
    ser = bench.make_packages(wl, 11, 5, workers=1)
    assert len(par) == len(ser) == 5
    for p, q in zip(par, ser):
        assert np.array_equal(p.cloud, q.cloud) and np.array_equal(p.imus, q.imus) and p.t0 == q.t0 and p.t1 == q.t1


def test_algorithmic_bytes_model():
    """bench.py's algorithmic bytes are SURVEY.md §8(d), strictly: 132 B per point per executed iteration, 84 B per world point,
    and the map formula split over the kernels that carry each term (the per-kernel terms add up to the formula)."""
    import bench
    st = dict(pt_iters=3 * 1000, n_ins=100, n_touch=50, refit_points=400, n_refit=10, n_merge=2, n_mergevox=7, n_full=300, n_mergeprobe=250)
    assert bench.algo_bytes("k_iekf_loop", st, 1000) == 132 * 3000
    assert bench.algo_bytes("k_set_scan", st, 1000) == 84 * 1000
    assert bench.algo_bytes("k_world_points", st, 1000) == 84 * 1000
    assert bench.algo_bytes("k_fill", st, 1000) == 144 * 100 + 152 * 50 + 72 * 400 + 432 * 10
    assert bench.algo_bytes("k_merge_rounds", st, 1000) == 192 * 7 + 672 * 2          # distinct (voxel, scan) probes, not merge() calls
    assert bench.algo_bytes("k_no_such_kernel", st, 1000) == 0
    assert bench.algo_bytes("k_seg_fill", st, 1000) == 0                              # bookkeeping of the implementation: no compulsory bytes
    total = 144 * 100 + 160 * 50 + 72 * 400 + 432 * 10 + 32 * 300 + 192 * 7 + 672 * 2
    assert bench.map_bytes(st) == total
    assert sum(bench.algo_bytes(k, st, 1000) for k in ("k_map_insert", "k_fill", "k_merge_rounds")) == total
    assert bench.algo_bytes("k_world_insert_count", st, 1000) == 84 * 1000 + bench.algo_bytes("k_map_insert", st, 1000)
    from voxelmapplus_fastlio2_b200.ctypes_defs import map_update_bytes
    assert map_update_bytes(st) == total
    assert bench.config_dict(bench.WORKLOADS["c2"], 40) == bench.config_dict(bench.WORKLOADS["c2"], 40)
    assert bench.DEFAULT_WORKLOAD == "c2" and bench.WORKLOADS["c2"]["pts"] == 200000 and bench.WORKLOADS["c2"]["max_iter"] == 4


def test_bench_stdout_carries_only_the_json_line():
    """bench.py's contract is ONE JSON line on stdout; whatever libraries print there (NCCL's version banner under torchrun) is sent to
    stderr by claim_stdout()."""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); print('library noise'); os.write(1, b'raw noise\\n'); "
            "bench.emit_line({'metric': 'scans_per_s', 'value': 1.0})" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    assert r.stdout == '{"metric": "scans_per_s", "value": 1.0}\n'
    assert "library noise" in r.stderr and "raw noise" in r.stderr
