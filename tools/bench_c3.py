"""C3 (BASELINE.json configs[2]): long city-block drive with LRU eviction continuously active and plane merging.

  python tools/bench_c3.py [--scans 10000] [--pts 20000] [--capacity 100000] [--out profiles/r01_c3_city.json]

Scene B (Manhattan grid of facades + ground, 200 m blocks; `scene_city(pilasters=True)`: shallow pilasters on the facades of the
driven street), back and forth along one street at 5 m/s mean speed (10 000 scans at 10 Hz = 5 km), 20 000 pts/scan, 0.5 m
voxels, map_capacity 100 000.  The whole LIOBuilder::process loop runs (host IMU propagation, device motion compensation +
IEKF + map update, one graph per scan); reported: scans/s of the loop, device time per scan (p50/p95), evictions / merges and
the drift against the ground-truth trajectory.

Why the pilasters: between cross streets the bare scene is two parallel facades and the ground; the along-street direction is
then unobservable for the reference's point-to-plane matcher, the estimator (CPU oracle and device path alike) picks up a
spurious velocity in the first scans, never loses it, and ends hundreds of kilometres away.  With facade relief the same
estimator stays within a few metres of the ground truth over the 5 km (0.06 %), and the CPU oracle shows the same error to the
millimetre on the scans compared (0.099 m at scan 500, 0.47 m at scan 1000).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voxelmapplus_fastlio2_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=10000)
    ap.add_argument("--pts", type=int, default=20000)
    ap.add_argument("--capacity", type=int, default=100000)
    ap.add_argument("--pipelined", type=int, default=0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--profile-from", type=int, default=-1, help="per-kernel device times (profiling mode: kernels launched one by one) from this scan on")
    a = ap.parse_args()
    # 2240 m per period, mean speed 5 m/s (peak 7.9 m/s)
    traj = synth.Trajectory(centre=(600.0, 600.0, 1.8), ax=560.0, ay=2.0, period=448.0)   # mid-street (facades at y = 588 and 612)
    seq = synth.Sequence(scene=synth.scene_city(pilasters=True), traj=traj, sensor=synth.SensorConfig(pts_per_scan=a.pts), seed=0xC3, cull=True)
    t = time.time()
    import multiprocessing as mp
    workers = max(1, min(16, os.cpu_count() or 1))
    with mp.get_context("fork").Pool(workers) as pool:           # before CUDA is touched
        clouds = dict(zip(range(a.scans), pool.map(seq.cloud, range(a.scans), chunksize=8)))
    pkgs = list(seq.packages(a.scans, clouds=clouds))
    del clouds
    print(f"[c3] generated {a.scans} packages in {time.time() - t:.1f}s ({workers} workers)", file=sys.stderr)

    from voxelmapplus_fastlio2_b200.ctypes_defs import default_config
    from voxelmapplus_fastlio2_b200.lio import LIOBuilder
    cfg = default_config(max_points_per_scan=a.pts + 64, map_capacity=a.capacity)
    lio = LIOBuilder(cfg, pipelined=bool(a.pipelined))
    gpu_ms, iters = [], []
    tot = dict(n_evicted=0, n_merge=0, n_created=0, n_refit=0, n_points=0)
    err = []
    path = 0.0
    last_gt = None
    t0 = None
    align = None
    failed = None
    for pk in pkgs:
        if a.profile_from >= 0 and pk.index == a.profile_from:
            lio.map.sync()
            lio.map.profile_enable(True)
            lio.map.profile_reset()
            n_prof0 = len(gpu_ms)
        try:
            st = lio.process(pk.imus, pk.cloud, pk.t0, pk.t1)
        except Exception as e:          # a capacity condition reported by the device: record where, keep what was measured
            failed = {"scan": pk.index, "error": str(e)}
            print(f"[c3] scan {pk.index}: {e}", file=sys.stderr)
            break
        if st.iters == 0:
            continue
        if t0 is None:
            lio.map.sync()
            t0, k0 = time.perf_counter(), pk.index
            # the estimator's world frame is gravity-aligned with yaw 0 at start-up: align it to the ground truth once
            x, _, _ = lio.state()
            Rg, Re = pk.gt_rot, np.array(x.rot[:]).reshape(3, 3)
            align = (Rg @ Re.T, np.array(x.pos[:]), pk.gt_pos.copy())
            continue
        gpu_ms.append(st.gpu_ms)
        iters.append(st.iters)
        for f in tot:
            tot[f] += getattr(st.map, f)
        if last_gt is not None:
            path += float(np.linalg.norm(pk.gt_pos - last_gt))
        last_gt = pk.gt_pos
        if pk.index % 500 == 0:
            x, _, _ = lio.state()
            e = float(np.linalg.norm(align[0] @ (np.array(x.pos[:]) - align[1]) - (pk.gt_pos - align[2])))
            err.append((pk.index, round(path, 1), round(e, 3), int(st.map.map_size)))
            print(f"[c3] scan {pk.index} path {path:.0f} m  |pos - gt| {e:.3f} m  map {st.map.map_size}  evicted so far {tot['n_evicted']}", file=sys.stderr)
    if failed is None:
        lio.map.sync()
    wall = time.perf_counter() - t0
    n = len(gpu_ms)
    x, _, _ = lio.state()
    res = {"config": "C3 city drive (BASELINE.json configs[2])", "scans": n, "pts_per_scan": a.pts, "voxel_size": 0.5, "map_capacity": a.capacity,
           "pipelined": bool(a.pipelined), "path_m": round(path, 1),
           "loop_scans_per_s": round(n / wall, 1), "device_ms_per_scan_p50": round(float(np.median(gpu_ms)), 4),
           "device_ms_per_scan_p95": round(float(np.percentile(gpu_ms, 95)), 4), "iters_mean": round(float(np.mean(iters)), 3),
           "totals": {k: int(v) for k, v in tot.items()}, "final_map_size": int(lio.map.map_size()) if failed is None else None, "failed": failed,
           "drift_vs_ground_truth": [{"scan": i, "path_m": p, "err_m": e, "map_size": m} for i, p, e, m in err],
           "note": "device time = CUDA events around upload + graph of vmp_scan_raw (motion compensation, IEKF, map update); "
                   "the loop adds host IMU propagation and synthetic-package handling; err_m = |estimated - ground-truth position| "
                   "after aligning the estimator's gravity-aligned start frame to the ground truth"}
    if a.profile_from >= 0:
        prof = lio.map.profile_read()
        k = max(1, n - n_prof0)
        res["per_kernel_us_per_scan"] = {name: round(ms * 1e3 / k, 2) for name, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]) if cnt > 0}
        res["profiled_scans"] = k
    txt = json.dumps(res, indent=1)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt)


if __name__ == "__main__":
    main()
