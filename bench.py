#!/usr/bin/env python
"""bench.py — scans/s of the voxel_plus per-scan hot path (IEKF measurement update + voxel-map update).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (CPU oracle on the host cores, rank 0 only)

A "step" is one scan through the timed region lio_builder.cpp:224-246 (calcBodyCov, IESKF::update with all
iterations, world points + covariances, VoxelMap::update).  IMU propagation / undistortion run on the host
between steps and are outside the timed region on both arms.

  value   whole-job scans/s with the scan and the prior already resident in HBM (vmp_scan_dev), timed per
          step with CUDA events on the launching stream, max over ranks
  e2e     the same through the host-buffer C-ABI: the host LIOBuilder writes the scan into the handle's PINNED staging
          area (vmp_scan_buffer) and calls vmp_scan_staged; the timed region (wall clock inside that call) holds the
          H2D copy of header + prior + points, the graph, and the posterior / counters written back to mapped host memory
  N > 1   replicas only: one independent synthetic sequence (own map) per GPU, no collective on the path
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case (and the one the >= 20x target is quoted on):
    # reported as the "c1" object of the default line, or alone with --workload c1
    "c1": dict(name="C1: synthetic Mid-360 sequence, 20k pts/scan @10 Hz, 0.5 m voxels, scene A (40x30x6 m hall)",
               pts=20000, voxel_size=0.5, max_iter=5, capacity=100000),
    # BASELINE.json configs[1]: the default workload of the bench line (largest single-GPU scan configuration)
    "c2": dict(name="C2: dense synthetic sequence, 200k pts/scan (no downsampling), 0.25 m voxels, 4 IEKF iterations, scene A",
               pts=200000, voxel_size=0.25, max_iter=4, capacity=400000),
}
DEFAULT_WORKLOAD = "c2"
C1_SIDE_STEPS = 100       # steps of the C1 side measurement that rides along in the default (C2) line


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md).

    One streaming nvidia-smi (-lms 20) is started EARLY (its start-up takes longer than the ~50 ms timed region of the
    default run); every sample is time-stamped and only those inside [mark_begin, stop] are reported.  If the window is
    shorter than the sampling period the nearest sample on either side (GPU busy with the same steps) is used and flagged."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []          # (t, sm, sm_max, reasons)
        self._halt = threading.Event()
        self.t_begin = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        for line in self.proc.stdout:
            out = line.strip().split(",")
            try:
                rs = [n for n, v in zip(names, out[2:]) if v.strip().lower().startswith("active")]
                self.samples.append((time.perf_counter(), float(out[0]), float(out[1]), rs))
            except Exception:
                pass
            if self._halt.is_set():
                break

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def stop(self):
        t_end = time.perf_counter()
        time.sleep(0.03)           # let the sample that covers the end of the window arrive
        self._halt.set()
        if getattr(self, "proc", None) is not None:
            self.proc.terminate()
        self.join(timeout=5)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        t0 = self.t_begin if self.t_begin is not None else self.samples[0][0]
        inside = [x for x in self.samples if t0 <= x[0] <= t_end]
        note = None
        if not inside:
            mid = 0.5 * (t0 + t_end)
            inside = sorted(self.samples, key=lambda x: abs(x[0] - mid))[:2]
            note = "timed region shorter than the 20 ms sampling period: nearest samples (same steps running)"
        sm = sorted(x[1] for x in inside)
        out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": inside[0][2], "reasons": sorted({r for x in inside for r in x[3]}),
               "samples": len(inside)}
        if note:
            out["note"] = note
        return out


def make_packages(wl, seed, count, workers=None):
    """Synthetic SyncPackages 0..count-1.  A scan is a pure function of (seed, index), so the ray casting is spread over
    the host cores (forked workers, numpy only; called before this process touches CUDA)."""
    from voxelmapplus_fastlio2_b200 import synth
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=wl["pts"]), seed=seed)
    t = time.time()
    if workers is None:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        workers = max(1, min(16, (os.cpu_count() or 1) // max(1, world)))
    clouds = None
    if workers > 1 and count >= 4:
        import multiprocessing as mp
        try:
            with mp.get_context("fork").Pool(workers) as pool:
                clouds = dict(zip(range(count), pool.map(seq.cloud, range(count), chunksize=1)))
        except Exception as e:       # no fork / no semaphores: generate serially
            log(f"[bench] parallel generation unavailable ({e}); serial")
            clouds = None
    pk = list(seq.packages(count, clouds=clouds))
    log(f"[bench] generated {count} synthetic packages of {wl['pts']} pts in {time.time() - t:.1f}s ({workers} workers)")
    return pk


def make_cfg(wl, device=0):
    from voxelmapplus_fastlio2_b200.ctypes_defs import default_config
    return default_config(max_points_per_scan=wl["pts"] + 64, voxel_size=wl["voxel_size"], opti_max_iter=wl["max_iter"],
                          map_capacity=wl["capacity"], device=device)


# algorithmic bytes per kernel class, from the per-scan counters
def algo_bytes(kernel, st_sum, n_pts_sum, iters_sum):
    if kernel == "k_measure":
        return 132 * st_sum["pt_iters"]
    if kernel == "k_set_scan":
        return (12 + 24 + 72) * n_pts_sum
    if kernel == "k_world_points":
        return 84 * n_pts_sum + (24 + 72) * n_pts_sum
    if kernel == "k_map_insert":
        return (24 + 8 + 4) * n_pts_sum
    if kernel == "k_map_count":
        return (4 + 4 + 12) * n_pts_sum
    if kernel == "k_seg_fill":
        return (4 + 4 + 4) * n_pts_sum
    if kernel == "k_log_append":
        return 4 * n_pts_sum + 20 * st_sum["n_touch"]
    if kernel == "k_fill_state":
        return 144 * st_sum["n_ins"] + 160 * st_sum["n_touch"] + 4 * n_pts_sum
    if kernel == "k_fill_refit":
        return 72 * st_sum["refit_points"] + 432 * st_sum["n_refit"]
    if kernel == "k_fill_acc":
        return 288 * st_sum["refit_points"] + 288 * st_sum["n_refit"]
    if kernel in ("k_merge_prefilter", "k_merge_rounds"):
        return 192 * st_sum["n_merge_voxels"] + 672 * st_sum["n_merge"]
    return 0


def run_ours(args):
    # synthetic input first: the generator forks worker processes and must run before this process touches CUDA
    rank0 = int(os.environ.get("RANK", "0"))
    seed = 0xC0FFEE + rank0                 # == replicas.rank_seed(rank)
    W, K = args.warmup, args.steps
    wl = WORKLOADS[args.workload]
    pkgs = make_packages(wl, seed, W + K + 2)
    side = None
    if args.workload != "c1" and not args.no_c1:
        Ks = min(K, C1_SIDE_STEPS) if args.c1_steps is None else args.c1_steps
        side = (WORKLOADS["c1"], Ks, make_packages(WORKLOADS["c1"], seed, W + Ks + 2))

    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    from voxelmapplus_fastlio2_b200 import replicas

    rank, world, local = replicas.rank_env()
    assert replicas.rank_seed(rank) == seed
    if world != args.gpus and world > 1:
        log(f"[bench] warning: WORLD_SIZE={world} but --gpus {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    line = measure_workload(args, wl, pkgs, W, K, rank, world, local, full=True)
    if side is not None:
        wl1, K1, pk1 = side
        l1 = measure_workload(args, wl1, pk1, W, K1, rank, world, local, full=False)
        if rank == 0:
            line["c1"] = {"what": "the same measurement on BASELINE.json configs[0] (the >= 20x target is quoted on it), "
                                  "riding along in the default line; alone: --workload c1",
                          "workload": wl1["name"], "steps": K1, "warmup": W, "value": l1["value"], "unit": "scans/s",
                          "ms_per_step": l1["ms_per_step"], "p50_ms": l1["p50_ms"], "p95_ms": l1["p95_ms"],
                          "iters_mean": l1["iters_mean"], "e2e": l1["e2e"], "cpu_baseline": l1["cpu_baseline"],
                          "gpu_launches": l1["gpu_launches"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_workload(args, wl, pkgs, W, K, rank, world, local, full=True):
    """One workload through the three passes (end to end, resident, per-kernel).  full=False skips the whole-host-loop
    pass and the per-kernel roofline (the C1 side measurement)."""
    import torch
    import torch.distributed as dist

    from voxelmapplus_fastlio2_b200 import replicas
    from voxelmapplus_fastlio2_b200.bindings import HotPath
    from voxelmapplus_fastlio2_b200.lio import LIOBuilder

    cfg = make_cfg(wl, device=local)

    # ---------------- pass 1: end to end through the host-buffer API (host LIOBuilder -> vmp_scan)
    lio = LIOBuilder(cfg, device_undistort=False)      # the timed region is exactly lio_builder.cpp:224-246 = vmp_scan
    clouds, priors, e2e_host_ms, e2e_gpu_ms, e2e_stats = [], [], [], [], []
    first = None
    launches0 = None
    for pk in pkgs:
        cloud = pk.cloud.copy()
        st = lio.process(pk.imus, cloud, pk.t0, pk.t1)
        x, P, status = lio.state()
        if status < 2 and st.map.n_points == 0:
            continue
        x0, P0 = lio.prior()
        xyz = np.ascontiguousarray(cloud[:, :3])
        if st.iters == 0:
            first = (x0, P0, xyz)
            continue
        clouds.append(xyz)
        priors.append(np.concatenate([np.frombuffer(bytes(x0), np.float64), P0.ravel()]))
        e2e_host_ms.append(st.host_ms)
        e2e_gpu_ms.append(st.gpu_ms)
        e2e_stats.append((st.iters, sum(st.effect_num[:st.iters]), st.map.as_dict() if hasattr(st.map, "as_dict") else None))
    n_scans = len(clouds)
    assert n_scans >= W + K, (n_scans, W, K)
    x_e2e, _, _ = lio.state()
    lio.close()
    h2d = int(4608 + clouds[W].nbytes)           # ScanIn header (prior included) + points, one DMA copy
    d2h = int(565 * 8 + 56 + 144)                # StateOut + MapOut mailboxes, written by the kernels

    # ---------------- pass 1b: the whole host loop (IMU propagation + undistortion + vmp_scan), synchronous vs pipelined
    def lio_loop(pipelined, device_undistort=True):
        lb = LIOBuilder(cfg, pipelined=pipelined, device_undistort=device_undistort)
        cl = [pk.cloud.copy() for pk in pkgs]
        t0 = None
        done = 0
        for k, pk in enumerate(pkgs):
            st = lb.process(pk.imus, cl[k], pk.t0, pk.t1)
            if t0 is None and st.iters > 0:
                done += 1
                if done == W:
                    lb.map.sync()
                    t0, k0 = time.perf_counter(), k
        lb.map.sync()
        dt = time.perf_counter() - t0
        n = len(pkgs) - 1 - k0
        lb.close()
        return n / dt
    sampler = ClockSampler(local)
    sampler.start()                      # streaming from here on; the reported window opens at step W of pass 2
    loop_host = loop_sync = loop_pipe = None
    if full and not args.no_loops:
        loop_host = lio_loop(False, device_undistort=False)
        loop_sync = lio_loop(False)
        loop_pipe = lio_loop(True)

    # ---------------- pass 2: resident replay (scan + prior already in HBM), timed per step with CUDA events
    g = HotPath(cfg)
    g.first_scan(*first)
    dev = torch.device("cuda", local)
    d_clouds = [torch.from_numpy(c).to(dev) for c in clouds]
    d_priors = [torch.from_numpy(p).to(dev) for p in priors]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    res_ms, res_stats = [], []
    launches_timed = 0
    for i in range(n_scans):
        if i == W:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            sampler.mark_begin()
            launches0 = g.launch_count()
            t_wall0 = time.perf_counter()
        if i == W + K:
            break
        flush.zero_()                       # L2 flush between steps (256 MiB > 126 MB L2), outside the events
        torch.cuda.synchronize()
        st = g.scan_dev(d_clouds[i].data_ptr(), clouds[i].shape[0], d_priors[i].data_ptr())
        assert st.iters == e2e_stats[i][0] and sum(st.effect_num[:st.iters]) == e2e_stats[i][1], \
            f"replay diverged from the end-to-end pass at scan {i}"
        res_stats.append(st)
        if i >= W:
            res_ms.append(st.gpu_ms)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop()
    launches_timed = g.launch_count() - launches0
    # replicas only: whole-job value = scans of all ranks / MAX over ranks of the device time
    total_ms_max = replicas.max_over_ranks(float(np.sum(res_ms)), dev)
    e2e_ms_max = replicas.max_over_ranks(float(np.sum(e2e_host_ms[W:W + K])), dev)
    g.close()

    # ---------------- pass 3 (rank 0): per-kernel CUDA-event timing of the same steps -> live roofline
    roof = None
    if rank == 0 and full:
        gp = HotPath(cfg)
        gp.first_scan(*first)
        gp.profile_enable(True)
        agg = dict(pt_iters=0, n_ins=0, n_touch=0, refit_points=0, n_refit=0, n_merge=0, n_merge_voxels=0)
        n_pts_sum = iters_sum = 0
        for i in range(W + K):
            if i == W:
                gp.profile_reset()
            flush.zero_()
            torch.cuda.synchronize()
            st = gp.scan_dev(d_clouds[i].data_ptr(), clouds[i].shape[0], d_priors[i].data_ptr())
            if i >= W:
                n = clouds[i].shape[0]
                n_pts_sum += n
                iters_sum += st.iters
                agg["pt_iters"] += n * st.iters
                for f in ("n_ins", "n_touch", "refit_points", "n_refit", "n_merge"):
                    agg[f] += getattr(st.map, f)
                agg["n_merge_voxels"] += st.map.n_touch
        prof = gp.profile_read()
        gp.close()
        tot = sum(v[0] for v in prof.values())
        ranked = sorted(prof.items(), key=lambda kv: -kv[1][0])
        top, (top_ms, top_launches) = ranked[0]
        # launches that do work: early-exited IEKF iterations are counted with the executed ones
        eff_launches = iters_sum if top == "k_measure" else max(1, top_launches)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        bytes_total = algo_bytes(top, agg, n_pts_sum, iters_sum)
        achieved = bytes_total / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
        traffic = None
        try:        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this workload
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(args.workload, {}).get(top)
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": top, "achieved": round(achieved, 3), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 6), "traffic": traffic,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": round(bytes_total / eff_launches, 1),
                "avg_launch_us": round(top_ms * 1e3 / eff_launches, 3),
                "kernel_share_of_step": round(top_ms / tot, 4) if tot > 0 else None,
                "per_kernel_us_per_step": {k: round(v[0] * 1e3 / K, 2) for k, v in ranked if v[1] > 0}}

    # ---------------- CPU baseline (rank 0, N == 1): the oracle on a bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(wl, pkgs, min(W + K, args.cpu_scans), threads=1)

    if rank == 0:
        p50 = float(np.median(res_ms))
        line = {
            "metric": "scans_per_s", "value": round(world * K / (total_ms_max * 1e-3), 2), "unit": "scans/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(total_ms_max / K, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "pts_per_scan": wl["pts"], "voxel_size": wl["voxel_size"],
                       "max_iter": wl["max_iter"], "map_capacity": wl["capacity"], "parallelism": f"replicas x{world}",
                       "l2": "flushed between steps (256 MiB memset, outside the timed events)",
                       "timing": "per-step CUDA events on the launching stream, summed over K steps, max over ranks"},
            "p50_ms": round(p50, 4), "p95_ms": round(float(np.percentile(res_ms, 95)), 4),
            "iters_mean": round(float(np.mean([s.iters for s in res_stats[W:]])), 3),
            "effect_num_mean": round(float(np.mean([s.effect_num[s.iters - 1] for s in res_stats[W:]])), 1),
            "wall_ms_per_step_incl_flush": round((t_wall1 - t_wall0) * 1e3 / K, 4),
            "e2e": {"value": round(world * K / (e2e_ms_max * 1e-3), 2), "unit": "scans/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "p50_ms": round(float(np.median(e2e_host_ms[W:W + K])), 4),
                    "timing": "wall clock inside the synchronous vmp_scan_staged: one H2D copy (header + prior + points) from the pinned staging area the host LIOBuilder filled, graph, mailbox write-back to mapped host memory, sync"},
            "host_loop": None if loop_host is None else {
                          "host_undistort_scans_per_s": round(loop_host, 1), "sync_scans_per_s": round(loop_sync, 1),
                          "pipelined_scans_per_s": round(loop_pipe, 1),
                          "what": "wall clock of the whole LIOBuilder.process loop, lio_builder.cpp:65-246 (host IMU propagation, motion "
                                  "compensation, update): compensation on the host + vmp_scan / on the device in the scan's graph "
                                  "(vmp_scan_raw) / the same with vmp_set_pipelined"},
            "gpu_launches": int(launches_timed),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        return line
    return None


def cpu_baseline(wl, pkgs, n_scans, threads=1):
    """The CPU oracle (a restatement: the reference cannot be built here, kind 'port') on the same packages."""
    from oracle.oracle_py import Oracle
    cfg = make_cfg(wl)
    o = Oracle(cfg)
    o.set_threads(threads)
    ms = []
    t = time.time()
    for pk in pkgs:
        st = o.lio_process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        if st.iters > 0:
            ms.append(st.gpu_ms)        # CPU milliseconds of lio_builder.cpp:224-246 for the oracle
        if len(ms) >= n_scans or time.time() - t > 40:
            break
    skip = min(5, len(ms) // 4)
    use = ms[skip:]
    return {"value": round(len(use) / (sum(use) * 1e-3), 3), "unit": "scans/s", "cores": threads, "kind": "port",
            "sample": f"{len(use)} scans of the same sequence after {skip} warm-up scans (timed region lio_builder.cpp:224-246, "
                      f"-O3, asserts on, like the reference build), {os.cpu_count()} host cores present",
            "p50_ms": round(float(np.median(use)), 3)}


def run_reference(args):
    """--impl reference: the reference's CPU path for the same metric/config.  The reference itself cannot be
    compiled here (Eigen/Sophus/PCL/ROS absent) so this is the oracle port, with all the host threads the
    reference can use on this path (its one OpenMP loop, lio_builder.cpp:256-274)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as ge
    from oracle import oracle_py
    oracle_py.build()
    del ge
    wl = WORKLOADS[args.workload]
    W, K = args.warmup, args.steps
    K = min(K, args.cpu_scans)
    threads = os.cpu_count() or 1
    pkgs = make_packages(wl, 0xC0FFEE, W + K + 2)
    from oracle.oracle_py import Oracle
    o = Oracle(make_cfg(wl))
    o.set_threads(threads)
    ms = []
    for pk in pkgs:
        st = o.lio_process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        if st.iters > 0:
            ms.append(st.gpu_ms)
    use = ms[W:W + K]
    v = round(len(use) / (sum(use) * 1e-3), 3)
    line = {"impl": "reference", "metric": "scans_per_s", "value": v, "unit": "scans/s", "n_gpus": args.gpus, "steps": len(use),
            "warmup": W, "ms_per_step": round(float(np.mean(use)), 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "pts_per_scan": wl["pts"], "voxel_size": wl["voxel_size"],
                       "max_iter": wl["max_iter"], "map_capacity": wl["capacity"]},
            "cpu_baseline": {"value": v, "unit": "scans/s", "cores": threads, "kind": "port",
                             "sample": f"{len(use)} scans after {W} warm-up scans; OpenMP on the one loop the reference "
                                       f"parallelises; the rest of the path is serial by construction"},
            "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-c1", action="store_true", help="skip the C1 side measurement of the default line")
    ap.add_argument("--c1-steps", type=int, default=None)
    ap.add_argument("--cpu-scans", type=int, default=300, help="bound of the CPU sample (scans)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-loops", action="store_true", help="skip the whole-host-loop pass (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
