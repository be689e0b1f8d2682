#!/usr/bin/env python
"""bench.py — scans/s of the voxel_plus per-scan hot path (IEKF measurement update + voxel-map update).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (CPU oracle on the host cores, rank 0 only)

A "step" is one scan through the timed region lio_builder.cpp:224-246 (calcBodyCov, IESKF::update with all
iterations, world points + covariances, VoxelMap::update).  IMU propagation / undistortion run on the host
between steps and are outside the timed region on both arms.

  value   whole-job scans/s with the scan and the prior already resident in HBM (vmp_scan_dev), timed per
          step with CUDA events on the launching stream, max over ranks
  e2e     the same through the host-buffer C-ABI, vmp_scan(handle, x, P, <plain pageable host pointer>, n): wall clock
          inside the call = copy of the caller's points into the pinned staging, ONE H2D copy (header + prior + points),
          the graph, posterior / counters written back to mapped host memory, sync
  e2e_lio the host LIOBuilder::process's timed region (lio_builder.cpp:224-246): reading x y z out of the caller's
          pageable cloud into the pinned staging (vmp_scan_buffer) + vmp_scan_staged
  N > 1   replicas only: one independent synthetic sequence (own map) per GPU, no collective on the path

The timed window is a MOVING sensor with a grown map: `--prime` (default 40) scans run untimed first (the synthetic
trajectory stands still for its first 2 s = 20 scans), then W warm-up scans, then the K timed ones.
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
PRIME_SCANS = 40          # untimed scans in front of warm-up + timed window (static start-up of the trajectory: 20 scans)
C1_SIDE_STEPS = 100       # steps of the C1 side measurement that rides along in the default (C2) line


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md), in-process through NVML
    (nvidia_ml_py) every ~1 ms, so that even a 10 ms window holds several samples; falls back to a streaming
    `nvidia-smi -lms 20` when NVML cannot be loaded.  Samples are time-stamped; only those inside
    [mark_begin, stop] are reported."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []          # (t, sm, sm_max, reasons)
        self._halt = threading.Event()
        self.t_begin = None
        self.source = None
        self.proc = None

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.index
        if vis:
            try:
                idx = int(vis.split(",")[self.index])
            except Exception:
                pass
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        self.source = "nvml"
        while not self._halt.is_set():
            sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            try:
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            self.samples.append((time.perf_counter(), sm, sm_max, [n for n, b in self.REASONS if bits & b]))
            time.sleep(0.001)

    def _run_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                      "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        self.source = "nvidia-smi -lms 20"
        for line in self.proc.stdout:
            out = line.strip().split(",")
            try:
                rs = [n for n, v in zip(names, out[2:]) if v.strip().lower().startswith("active")]
                self.samples.append((time.perf_counter(), float(out[0]), float(out[1]), rs))
            except Exception:
                pass
            if self._halt.is_set():
                break

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            try:
                self._run_smi()
            except Exception:
                pass

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def stop(self):
        t_end = time.perf_counter()
        time.sleep(0.03)           # let the sample that covers the end of the window arrive
        self._halt.set()
        if self.proc is not None:
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
            note = "timed region shorter than the sampling period: nearest samples (same steps running)"
        sm = sorted(x[1] for x in inside)
        out = {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": inside[0][2],
               "reasons": sorted({r for x in inside for r in x[3]}), "samples": len(inside), "source": self.source}
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


def config_dict(wl, prime):
    """`config` of the JSON line: identical on both arms (repo / --impl reference)."""
    return {"workload": wl["name"], "pts_per_scan": wl["pts"], "voxel_size": wl["voxel_size"], "max_iter": wl["max_iter"],
            "map_capacity": wl["capacity"], "prime_scans": prime}


# ALGORITHMIC bytes, strictly SURVEY.md §8(d) (compulsory traffic, layout-independent; counters come from the run itself):
#   IEKF measurement   132 B per point per executed iteration (24 point_lidar + 48 cov_lidar + 8 hash key + 52 plane)
#   world points       84 B per point (12 B float xyz in, 72 B pv out)
#   map update         144 N_ins + 160 N_touch + sum_refits(72 n + 432) + 32 N_full + 192 N_mergeprobe + 672 N_merge
#                      with N_mergeprobe = distinct (full plane voxel, scan) merge evaluations = the counter n_mergevox
# The map formula is split over the kernels that carry the respective term; kernels that only do the implementation's
# own bookkeeping (segment build, LRU log, finalize) have NO algorithmic bytes and only count in the map update's time.
def algo_bytes(kernel, st_sum, n_pts_sum):
    if kernel in ("k_iekf_loop", "k_measure", "k_iekf"):      # all IEKF iterations of a scan (one resident launch since round 2)
        return 132 * st_sum["pt_iters"]
    if kernel == "k_set_scan":
        return 84 * n_pts_sum                       # 12 B in, point_lidar 24 + cov_lidar 48 out
    if kernel == "k_world_points":
        return 84 * n_pts_sum
    if kernel == "k_map_insert":
        return 32 * st_sum["n_full"] + 8 * st_sum["n_touch"]          # points into full voxels (24 + 8 key); key of a touch
    if kernel == "k_world_insert_count":                              # one kernel since round 2: pv_list production + find-or-create
        return 84 * n_pts_sum + 32 * st_sum["n_full"] + 8 * st_sum["n_touch"]
    if kernel == "k_fill_heavy":
        return 0                                                      # (the CTA-path launch of k_fill: its bytes are counted with "k_fill")
    if kernel in ("k_fill", "k_fill_state"):
        # one kernel since round 2: append (72 in + 72 out), n / mean / ppt read + write, and the refits (stored points re-read, 6x6 cov + normal)
        return 144 * st_sum["n_ins"] + 152 * st_sum["n_touch"] + 72 * st_sum["refit_points"] + 432 * st_sum["n_refit"]
    if kernel in ("k_merge_prefilter", "k_merge_rounds"):
        return 192 * st_sum["n_mergevox"] + 672 * st_sum["n_merge"]
    return 0


MAP_KERNELS = ("k_world_insert_count", "k_map_begin", "k_map_insert", "k_map_count", "k_seg_scan", "k_seg_fill", "k_lru_evict", "k_fill", "k_fill_heavy", "k_fill_classify", "k_fill_state", "k_fill_refit",
               "k_fill_acc", "k_merge_prefilter", "k_merge_rounds", "k_log_append", "k_map_finalize", "k_map_end")


def map_bytes(st_sum):
    return (144 * st_sum["n_ins"] + 160 * st_sum["n_touch"] + 72 * st_sum["refit_points"] + 432 * st_sum["n_refit"]
            + 32 * st_sum["n_full"] + 192 * st_sum["n_mergevox"] + 672 * st_sum["n_merge"])


def run_ours(args):
    # synthetic input first: the generator forks worker processes and must run before this process touches CUDA
    rank0 = int(os.environ.get("RANK", "0"))
    seed = 0xC0FFEE + rank0                 # == replicas.rank_seed(rank)
    W, K, PR = args.warmup, args.steps, args.prime
    wl = WORKLOADS[args.workload]
    pkgs = make_packages(wl, seed, PR + W + K + 3)
    side = None
    if args.workload != "c1" and not args.no_c1:
        Ks = min(K, C1_SIDE_STEPS) if args.c1_steps is None else args.c1_steps
        side = (WORKLOADS["c1"], Ks, make_packages(WORKLOADS["c1"], seed, PR + W + Ks + 3))

    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    from voxelmapplus_fastlio2_b200 import replicas

    rank, world, local = replicas.rank_env()
    assert replicas.rank_seed(rank) == seed
    if world != args.gpus and world > 1:
        log(f"[bench] warning: WORLD_SIZE={world} but --gpus {args.gpus}")
    # the staging helpers of a handle (vmp_stage.hpp: up to 11 + the caller) share the host with the other ranks of the node
    os.environ.setdefault("VMP_COPY_THREADS", str(max(1, min(11, (os.cpu_count() or 4) // max(1, world) - 2))))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    line = measure_workload(args, wl, pkgs, W, K, PR, rank, world, local, full=True)
    if side is not None:
        wl1, K1, pk1 = side
        l1 = measure_workload(args, wl1, pk1, W, K1, PR, rank, world, local, full=False)
        if rank == 0:
            line["c1"] = {"what": "the same measurement on BASELINE.json configs[0] (the >= 20x target is quoted on it), "
                                  "riding along in the default line; alone: --workload c1",
                          "config": l1["config"], "steps": K1, "warmup": W, "value": l1["value"], "unit": "scans/s",
                          "ms_per_step": l1["ms_per_step"], "p50_ms": l1["p50_ms"], "p95_ms": l1["p95_ms"],
                          "iters_mean": l1["iters_mean"], "e2e": l1["e2e"], "e2e_lio": l1["e2e_lio"], "cpu_baseline": l1["cpu_baseline"],
                          "gpu_launches": l1["gpu_launches"]}
    if rank == 0:
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_workload(args, wl, pkgs, W, K, PR, rank, world, local, full=True):
    """One workload through the passes (end to end, resident, per-kernel).  full=False skips the whole-host-loop
    pass and the per-kernel roofline (the C1 side measurement).  The first PR + W scans of every pass are untimed."""
    import torch
    import torch.distributed as dist

    from voxelmapplus_fastlio2_b200 import replicas
    from voxelmapplus_fastlio2_b200.bindings import HotPath
    from voxelmapplus_fastlio2_b200.lio import LIOBuilder

    cfg = make_cfg(wl, device=local)
    dev = torch.device("cuda", local)
    S0 = PR + W                                         # first timed scan (index into the LIO_MAPPING scans)

    # ---------------- pass 1: the host LIOBuilder (lio_builder.cpp:175-248 with host compensation); its timed region is
    # lio_builder.cpp:224-246 incl. the read of x y z out of the caller's pageable cloud into the pinned staging
    lio = LIOBuilder(cfg, device_undistort=False)
    clouds, priors, lio_host_ms, e2e_stats = [], [], [], []
    first = None
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
        priors.append((x0, P0, np.concatenate([np.frombuffer(bytes(x0), np.float64), P0.ravel()])))
        lio_host_ms.append(st.host_ms)
        e2e_stats.append((st.iters, sum(st.effect_num[:st.iters])))
    n_scans = len(clouds)
    assert n_scans >= S0 + K, (n_scans, PR, W, K)
    lio.close()
    h2d = int(4608 + 8192 + clouds[S0].nbytes)   # ScanIn header (prior included) + pose block + points, one DMA copy
    d2h = int(565 * 8 + 56 + 8 + 12 * 8 + 48)    # StateOut + MapOut mailboxes, written by the kernels

    # ---------------- pass 1b: e2e through the plain C-ABI call, vmp_scan with a pageable host pointer
    g = HotPath(cfg)
    g.first_scan(*first)
    e2e_ms = []
    for i in range(S0 + K):
        if i == S0 and world > 1:
            dist.barrier()
        x0, P0, _ = priors[i]
        _, _, st = g.scan(x0, P0, clouds[i])
        assert st.iters == e2e_stats[i][0] and sum(st.effect_num[:st.iters]) == e2e_stats[i][1], f"vmp_scan diverged from LIOBuilder at scan {i}"
        e2e_ms.append(st.host_ms)
    g.close()

    # ---------------- pass 1c: the whole host loop (IMU propagation + compensation + update), per rank
    def lio_loop(pipelined, device_undistort=True, writeback=True, device_predict=False):
        lb = LIOBuilder(cfg, pipelined=pipelined, device_undistort=device_undistort, cloud_writeback=writeback, device_predict=device_predict)
        cl = [pk.cloud.copy() for pk in pkgs]
        t0 = None
        done = 0
        for k, pk in enumerate(pkgs):
            st = lb.process(pk.imus, cl[k], pk.t0, pk.t1)
            if t0 is None and st.iters > 0:
                done += 1
                if done == S0:
                    lb.map.sync()
                    if world > 1:
                        dist.barrier()
                    t0, k0 = time.perf_counter(), k
        lb.map.sync()
        dt = time.perf_counter() - t0
        n = len(pkgs) - 1 - k0
        lb.close()
        return world * n / replicas.max_over_ranks(dt, dev)
    sampler = ClockSampler(local)
    sampler.start()                      # streaming from here on; the reported window opens at the first timed step of pass 2
    loop_host = loop_sync = loop_pipe = loop_sync_lazy = loop_pipe_lazy = loop_sync_pred = loop_pipe_pred = None
    if full and not args.no_loops:
        loop_host = lio_loop(False, device_undistort=False)
        loop_sync = lio_loop(False)
        loop_pipe = lio_loop(True)
        loop_sync_lazy = lio_loop(False, writeback=False)
        loop_pipe_lazy = lio_loop(True, writeback=False)
        loop_sync_pred = lio_loop(False, device_predict=True)
        loop_pipe_pred = lio_loop(True, device_predict=True)

    # ---------------- pass 2: resident replay (scan + prior already in HBM), timed per step with CUDA events
    g = HotPath(cfg)
    g.first_scan(*first)
    d_clouds = [torch.from_numpy(c).to(dev) for c in clouds[:S0 + K]]
    d_priors = [torch.from_numpy(p[2]).to(dev) for p in priors[:S0 + K]]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    res_ms, res_stats = [], []
    launches0 = 0
    for i in range(S0 + K):
        if i == S0:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            sampler.mark_begin()
            launches0 = g.launch_count()
            t_wall0 = time.perf_counter()
            if args.ncu_range:              # ncu --profile-from-start off: only the timed steps of this pass are captured
                torch.cuda.cudart().cudaProfilerStart()
        if not args.no_l2_flush:
            flush.zero_()                   # L2 flush between steps (256 MiB > 126 MB L2), outside the events
        torch.cuda.synchronize()
        st = g.scan_dev(d_clouds[i].data_ptr(), clouds[i].shape[0], d_priors[i].data_ptr())
        assert st.iters == e2e_stats[i][0] and sum(st.effect_num[:st.iters]) == e2e_stats[i][1], \
            f"replay diverged from the end-to-end pass at scan {i}"
        if i >= S0:
            res_ms.append(st.gpu_ms)
            res_stats.append(st)
    torch.cuda.synchronize()
    if args.ncu_range:
        torch.cuda.cudart().cudaProfilerStop()
    if world > 1:
        dist.barrier()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop()
    launches_timed = g.launch_count() - launches0
    # replicas only: whole-job value = scans of all ranks / MAX over ranks of the device time
    total_ms_max = replicas.max_over_ranks(float(np.sum(res_ms)), dev)
    e2e_ms_max = replicas.max_over_ranks(float(np.sum(e2e_ms[S0:S0 + K])), dev)
    lio_ms_max = replicas.max_over_ranks(float(np.sum(lio_host_ms[S0:S0 + K])), dev)
    g.close()

    # ---------------- pass 3 (rank 0): per-kernel CUDA-event timing of the same steps -> live roofline
    roof = None
    if rank == 0 and full and not args.ncu_range:
        gp = HotPath(cfg)
        gp.first_scan(*first)
        gp.profile_enable(True)
        agg = dict(pt_iters=0, n_ins=0, n_touch=0, refit_points=0, n_refit=0, n_merge=0, n_mergevox=0, n_full=0, n_mergeprobe=0)
        n_pts_sum = iters_sum = 0
        for i in range(S0 + K):
            if i == S0:
                gp.profile_reset()
            if not args.no_l2_flush:
                flush.zero_()
            torch.cuda.synchronize()
            st = gp.scan_dev(d_clouds[i].data_ptr(), clouds[i].shape[0], d_priors[i].data_ptr())
            if i >= S0:
                n = clouds[i].shape[0]
                n_pts_sum += n
                iters_sum += st.iters
                agg["pt_iters"] += n * st.iters
                for f in ("n_ins", "n_touch", "refit_points", "n_refit", "n_merge", "n_mergevox", "n_full", "n_mergeprobe"):
                    agg[f] += getattr(st.map, f)
        prof = gp.profile_read()
        gp.close()
        tot = sum(v[0] for v in prof.values())
        ranked = sorted(prof.items(), key=lambda kv: -kv[1][0])
        top, (top_ms, top_launches) = ranked[0]
        # launches that do work: early-exited IEKF iterations are counted with the executed ones
        eff_launches = max(1, top_launches)         # (k_iekf_loop: one launch per scan, all its iterations inside)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        bytes_total = algo_bytes(top, agg, n_pts_sum)
        achieved = bytes_total / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
        traffic = None
        try:        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this workload
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(args.workload, {}).get(top)
        except Exception:
            pass
        per_kernel = {}
        for k, v in ranked:
            if v[1] <= 0:
                continue
            b = algo_bytes(k, agg, n_pts_sum)
            per_kernel[k] = {"us_per_step": round(v[0] * 1e3 / K, 2), "algorithmic_MB_per_step": round(b / K / 1e6, 3),
                             "GBps": round(b / (v[0] * 1e-3) / 1e9, 1) if v[0] > 0 else None,
                             "frac": round(b / (v[0] * 1e-3) / 1e9 / peak, 5) if v[0] > 0 else None}
        map_ms = sum(v[0] for k, v in prof.items() if k in MAP_KERNELS)
        mb = map_bytes(agg) + 84 * n_pts_sum              # + the pv_list production (84 B / point), which opens the map update (k_world_insert_count)
        iekf_ms = sum(v[0] for k, v in prof.items() if k in ("k_set_scan", "k_iekf_loop", "k_measure", "k_iekf", "k_ieskf_solve", "k_scan_out"))
        ib = 132 * agg["pt_iters"] + 84 * n_pts_sum
        roof = {"bound": "hbm", "kernel": top, "achieved": round(achieved, 3), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 6), "traffic": traffic,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                "byte_model": "SURVEY.md 8(d), strictly: see algo_bytes() in bench.py",
                "algorithmic_bytes_per_launch": round(bytes_total / eff_launches, 1),
                "avg_launch_us": round(top_ms * 1e3 / eff_launches, 3),
                "kernel_share_of_step": round(top_ms / tot, 4) if tot > 0 else None,
                "map_update": {"what": "pv_list production (84 B / point) + whole VoxelMap::update: 8(d) map bytes / summed time of all map-update kernels",
                               "algorithmic_MB_per_step": round(mb / K / 1e6, 3), "us_per_step": round(map_ms * 1e3 / K, 2),
                               "achieved": round(mb / (map_ms * 1e-3) / 1e9, 2) if map_ms > 0 else None,
                               "frac": round(mb / (map_ms * 1e-3) / 1e9 / peak, 5) if map_ms > 0 else None,
                               "counters_per_step": {k: round(v / K, 1) for k, v in agg.items() if k != "pt_iters"}},
                "iekf": {"what": "set_scan + all measurement iterations: 132 B x point-iterations + 84 B x points",
                         "algorithmic_MB_per_step": round(ib / K / 1e6, 3), "us_per_step": round(iekf_ms * 1e3 / K, 2),
                         "achieved": round(ib / (iekf_ms * 1e-3) / 1e9, 2) if iekf_ms > 0 else None,
                         "frac": round(ib / (iekf_ms * 1e-3) / 1e9 / peak, 5) if iekf_ms > 0 else None},
                "per_kernel": per_kernel}

    # ---------------- CPU baseline (rank 0, N == 1): the oracle on a bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(wl, pkgs, S0, min(K, args.cpu_scans), threads=1)

    if rank == 0:
        p50 = float(np.median(res_ms))
        line = {
            "metric": "scans_per_s", "value": round(world * K / (total_ms_max * 1e-3), 2), "unit": "scans/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(total_ms_max / K, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(wl, PR),
            "protocol": {"parallelism": f"replicas x{world} (one independent sequence and map per GPU, no collective on the path)",
                         "l2": ("NOT flushed between steps (--no-l2-flush: diagnostic run, not a bench value)" if args.no_l2_flush
                                else "flushed between steps (256 MiB memset, outside the timed events)"),
                         "timing": "per-step CUDA events on the launching stream, summed over K steps, max over ranks",
                         "window": f"{PR} untimed priming scans (the trajectory stands still for its first 20) + {W} warm-up scans, "
                                   f"then {K} timed scans of a moving sensor"},
            "p50_ms": round(p50, 4), "p95_ms": round(float(np.percentile(res_ms, 95)), 4),
            "iters_mean": round(float(np.mean([s.iters for s in res_stats])), 3),
            "effect_num_mean": round(float(np.mean([s.effect_num[s.iters - 1] for s in res_stats])), 1),
            "wall_ms_per_step_incl_flush": round((t_wall1 - t_wall0) * 1e3 / K, 4),
            "e2e": {"value": round(world * K / (e2e_ms_max * 1e-3), 2), "unit": "scans/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "p50_ms": round(float(np.median(e2e_ms[S0:S0 + K])), 4),
                    "timing": "wall clock inside the synchronous vmp_scan(handle, x, P, pageable host pointer, n): copy of the caller's points into the pinned staging, "
                              "one H2D copy (header + prior + points), graph, mailbox write-back to mapped host memory, sync"},
            "e2e_lio": {"value": round(world * K / (lio_ms_max * 1e-3), 2), "unit": "scans/s",
                        "p50_ms": round(float(np.median(lio_host_ms[S0:S0 + K])), 4),
                        "timing": "the host LIOBuilder::process's timed region lio_builder.cpp:224-246: x y z read out of the caller's pageable cloud into "
                                  "the pinned staging (vmp_scan_buffer), then vmp_scan_staged"},
            "host_loop": None if loop_host is None else {
                          "host_undistort_scans_per_s": round(loop_host, 1), "sync_scans_per_s": round(loop_sync, 1),
                          "pipelined_scans_per_s": round(loop_pipe, 1),
                          "sync_no_writeback_scans_per_s": round(loop_sync_lazy, 1), "pipelined_no_writeback_scans_per_s": round(loop_pipe_lazy, 1),
                          "sync_device_predict_scans_per_s": round(loop_sync_pred, 1), "pipelined_device_predict_scans_per_s": round(loop_pipe_pred, 1),
                          "what": "wall clock of the whole LIOBuilder.process loop, lio_builder.cpp:65-246 (host IMU propagation, motion "
                                  "compensation, update), all ranks / max over ranks: compensation on the host + vmp_scan / on the device in the "
                                  "scan's graph (vmp_scan_raw) / the same with vmp_set_pipelined; no_writeback: the compensated cloud is not copied back into the "
                                  "caller's buffer (vmp_set_raw_writeback 0; it stays readable through vmp_get_lidar_cloud); device_predict: IESKF::predict on the device too "
                                  "(vmp_scan_raw_predict, SURVEY 8f row 2; A/B against the host propagation of the two lines above it)"},
            "gpu_launches": int(launches_timed),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        return line
    return None


def cpu_baseline(wl, pkgs, skip_scans, n_scans, threads=1, budget_s=75.0):
    """The CPU oracle (a restatement: the reference cannot be built here, kind 'port') on the same packages: the first
    `skip_scans` LIO scans run untimed (the same priming + warm-up as the device arm), the next n_scans are the sample."""
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
        if len(ms) >= skip_scans + n_scans or (len(ms) > skip_scans + 3 and time.time() - t > budget_s):
            break
    use = ms[skip_scans:]
    return {"value": round(len(use) / (sum(use) * 1e-3), 3), "unit": "scans/s", "cores": threads, "kind": "port",
            "sample": f"{len(use)} scans of the same sequence after {skip_scans} untimed scans (priming + warm-up, like the device arm); timed region "
                      f"lio_builder.cpp:224-246, -O3, asserts on, like the reference build; {os.cpu_count()} host cores present",
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
    W, K, PR = args.warmup, args.steps, args.prime
    K = min(K, args.cpu_scans)
    threads = os.cpu_count() or 1
    pkgs = make_packages(wl, 0xC0FFEE, PR + W + K + 3)
    cb = cpu_baseline(wl, pkgs, PR + W, K, threads=threads, budget_s=150.0)
    v = cb["value"]
    n_used = int(cb["sample"].split(" ")[0])
    line = {"impl": "reference", "metric": "scans_per_s", "value": v, "unit": "scans/s", "n_gpus": args.gpus, "steps": n_used,
            "warmup": W, "ms_per_step": round(1e3 / v, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(wl, PR),
            "cpu_baseline": {"value": v, "unit": "scans/s", "cores": threads, "kind": "port",
                             "sample": cb["sample"] + "; OpenMP on the one loop the reference parallelises, the rest of the path is serial by construction"},
            "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_line(line)


_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL: 'NCCL version ...' at communicator creation when
    NCCL_DEBUG is set in the environment), so the process's stdout is kept aside for the line and everything else goes to stderr."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit_line(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--prime", type=int, default=PRIME_SCANS, help="untimed scans in front of the warm-up (static start-up + map growth)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-c1", action="store_true", help="skip the C1 side measurement of the default line")
    ap.add_argument("--c1-steps", type=int, default=None)
    ap.add_argument("--cpu-scans", type=int, default=300, help="bound of the CPU sample (scans)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-loops", action="store_true", help="skip the whole-host-loop pass (profiling runs)")
    ap.add_argument("--no-l2-flush", action="store_true", help="diagnostic: leave the L2 as the previous scan left it (the default flushes it between steps)")
    ap.add_argument("--ncu-range", action="store_true", help="cudaProfilerStart/Stop around the timed steps of the resident pass (ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
