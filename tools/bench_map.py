"""C4 (BASELINE.json configs[3]): map-update microbench — world points + covariances through VoxelMap::build / update
with plane fit, merge and LRU eviction active, reported per kernel against the HBM roofline.

  python tools/bench_map.py [--points 10000000] [--batch 200000] [--capacity 100000] [--cpu-batches 6]

Scene B (Manhattan facades + ground), ground-truth poses (no filter): the sensor drives down a street, every batch is the
points it sees from its current position (range 40 m), float32-rounded like the reference's pv_list (Q15).
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
from voxelmapplus_fastlio2_b200.bindings import HotPath  # noqa: E402
from voxelmapplus_fastlio2_b200.ctypes_defs import default_config, map_update_bytes  # noqa: E402


def make_batches(n_batches, batch, seed=0xC4):
    scene = synth.scene_city()
    out = []
    t = time.time()
    for b in range(n_batches):
        rng = np.random.Generator(np.random.Philox(key=seed, counter=[b, 0, 0, 0]))
        # drive along the street between block rows 2 and 3, 3 m per batch, turning around at the ends
        s = (3.0 * b) % 2000.0
        x = 100.0 + (s if s < 1000.0 else 2000.0 - s)
        o = np.array([x, 600.0 + 1.5 * np.sin(0.05 * b), 1.8])
        az = rng.uniform(0, 2 * np.pi, batch)
        el = np.deg2rad(rng.uniform(-25, 50, batch))
        d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], 1)
        r = scene.cast(np.broadcast_to(o, d.shape).copy(), d, 0.5, 40.0)
        ok = np.isfinite(r)
        p = o + d[ok] * (r[ok] + rng.normal(0, 0.02, ok.sum()))[:, None]
        p = p.astype(np.float32).astype(np.float64)                      # Q15
        n = len(p)
        cov = np.zeros((n, 9))
        rr = r[ok]
        cov[:, 0] = cov[:, 4] = cov[:, 8] = 4e-4 + (1.75e-3 * rr) ** 2  # range + bearing noise, isotropic stand-in
        cov[:, 1] = 1e-7; cov[:, 3] = 1.0000001e-7
        out.append((p, cov))
    print(f"[bench_map] generated {n_batches} batches (~{batch} rays each) in {time.time() - t:.1f}s", file=sys.stderr)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--batch", type=int, default=200_000)
    ap.add_argument("--capacity", type=int, default=100_000)
    ap.add_argument("--cpu-batches", type=int, default=6)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    nb = max(3, a.points // a.batch)
    batches = make_batches(nb, a.batch)
    cfg = default_config(max_points_per_scan=a.batch + 64, map_capacity=a.capacity)
    g = HotPath(cfg)
    g.profile_enable(True)
    tot = dict(n_points=0, n_ins=0, n_touch=0, n_refit=0, refit_points=0, n_full=0, n_mergeprobe=0, n_merge=0, n_evicted=0)
    wall = 0.0
    warm = 2
    for k, (p, c) in enumerate(batches):
        if k == warm:
            g.profile_reset()
        t0 = time.perf_counter()
        st = g.map_update(p, c) if k else g.map_build(p, c)
        dt = time.perf_counter() - t0
        if k >= warm:
            wall += dt
            for f in tot:
                tot[f] += st[f]
    prof = g.profile_read()
    dev_ms = sum(v[0] for v in prof.values())
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    n = tot["n_points"]
    per_kernel_bytes = {
        "k_map_insert": 36 * n, "k_map_count": 20 * n, "k_seg_fill": 12 * n, "k_log_append": 4 * n + 20 * tot["n_touch"],
        "k_fill_state": 144 * tot["n_ins"] + 160 * tot["n_touch"] + 4 * n,
        "k_fill_refit": 72 * tot["refit_points"] + 432 * tot["n_refit"],
        "k_fill_acc": 288 * tot["refit_points"] + 288 * tot["n_refit"],
        "k_merge_prefilter": 192 * tot["n_touch"], "k_merge_rounds": 672 * tot["n_merge"],
    }
    rows = []
    for k, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        if cnt == 0:
            continue
        by = per_kernel_bytes.get(k, 0)
        gbs = by / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        rows.append({"kernel": k, "ms_total": round(ms, 3), "share": round(ms / dev_ms, 4), "algorithmic_GB": round(by / 1e9, 4),
                     "achieved_GBs": round(gbs, 2), "frac_of_measured_hbm": round(gbs / peak, 5)})
    algo = map_update_bytes(tot)
    res = {"config": "C4 map-update microbench", "points": n, "batches": nb - warm, "batch": a.batch, "capacity": a.capacity,
           "counters": tot, "device_ms_total": round(dev_ms, 2), "points_per_s_device": round(n / (dev_ms * 1e-3)),
           "points_per_s_host_api": round(n / wall), "h2d_bytes_per_point": 96,
           "algorithmic_bytes_total": algo, "achieved_GBs_whole_update": round(algo / (dev_ms * 1e-3) / 1e9, 2),
           "frac_of_measured_hbm_whole_update": round(algo / (dev_ms * 1e-3) / 1e9 / peak, 5), "hbm_peak_GBs": peak,
           "per_kernel": rows}
    if a.cpu_batches > 0:
        from oracle.oracle_py import Oracle
        o = Oracle(cfg)
        sec, pts = 0.0, 0
        for k, (p, c) in enumerate(batches[:a.cpu_batches + 1]):
            if k == 0:
                o.map_build(p, c)
                continue
            s, _ = o.map_update_timed(p, c)
            sec += s; pts += len(p)
        res["cpu_oracle_points_per_s"] = round(pts / sec)
        res["speedup_device_vs_cpu"] = round(res["points_per_s_device"] / res["cpu_oracle_points_per_s"], 1)
    txt = json.dumps(res, indent=1)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt)


if __name__ == "__main__":
    main()
