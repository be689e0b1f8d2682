"""Parity procedures at the BASELINE.json configuration sizes (C1 .. C4), shared by tests/test_gpu_parity_sizes.py (pytest,
`-m gpu`) and tools/parity_run.py (full lengths, JSON summary under profiles/).  Every procedure drives the CUDA path through
the C ABI and the CPU oracle with identical inputs, asserts the three tiers of BASELINE.json north_star and returns a
summary dict.

  tier 1  voxel keys, found / is_plane / is_valid per point, counters, evicted keys, map keys / flags / LRU order  bit-exact
  tier 1+ plane parameters of the map when both sides see identical world points                                  bit-exact
  tier 2  residuals, H, b                                                                                          <= 1e-9 relative
  tier 3  free-running trajectory                                                                                  <= 1 mm, 0.01 deg
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

from helpers import assert_maps_equal  # noqa: E402
from voxelmapplus_fastlio2_b200 import synth  # noqa: E402
from voxelmapplus_fastlio2_b200.ctypes_defs import default_config  # noqa: E402

RTOL_T2 = 1e-9
COUNTERS = ("n_ins", "n_touch", "n_created", "n_refit", "refit_points", "n_full", "n_mergeprobe", "n_mergevox", "n_merge", "n_evicted", "map_size")


def _packages(seq, count):
    """count packages of `seq`; the ray casting is spread over forked worker processes.  (The calling process usually holds a
    CUDA context already; the workers only run numpy and never touch it, like the workers of a data loader.)"""
    import multiprocessing as mp
    workers = max(1, min(16, (os.cpu_count() or 1)))
    clouds = None
    if workers > 1 and count >= 8:
        try:
            with mp.get_context("fork").Pool(workers) as pool:
                clouds = dict(zip(range(count), pool.map_async(seq.cloud, range(count), chunksize=4).get(timeout=900)))
        except Exception:
            clouds = None
    return list(seq.packages(count, clouds=clouds))


def _rot(x):
    return np.array(x.rot[:]).reshape(3, 3)


# ----------------------------------------------------------------------------------------------------------------- C1
def c1_free_running(oracle_mod, scans=300, pts=20000):
    """BASELINE configs[0] as specified: 20 000 pts/scan, 0.5 m voxels, reference defaults, free-running vmp::LIOBuilder (host IMU
    propagation, device motion compensation + IEKF + map update) against the free-running oracle.  Tier 3 over the whole run,
    executed iterations equal in every scan; effect_num is reported (the two runs see posteriors that differ in the 9th digit,
    so a gate decision at a margin below that may legitimately fall the other way: the test bounds how often)."""
    from voxelmapplus_fastlio2_b200.lio import LIOBuilder
    cfg = default_config(max_points_per_scan=pts + 64, voxel_size=0.5, map_capacity=100000)
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=pts))
    pk = _packages(seq, scans + 3)
    o = oracle_mod.Oracle(cfg)
    b = LIOBuilder(cfg)
    worst_p = worst_r = 0.0
    n_lio = n_eff_equal = 0
    worst_eff = 0
    t0 = time.time()
    for p in pk:
        so = o.lio_process(p.imus, p.cloud.copy(), p.t0, p.t1)
        sb = b.process(p.imus, p.cloud.copy(), p.t0, p.t1)
        xo, _, s1 = o.lio_state()
        xb, _, s2 = b.state()
        assert s1 == s2
        if s1 == 2 and so.iters:
            n_lio += 1
            assert sb.iters == so.iters, f"scan {p.index}: iterations {sb.iters} vs {so.iters}"
            d = max(abs(int(a) - int(c)) for a, c in zip(sb.effect_num[:so.iters], so.effect_num[:so.iters]))
            worst_eff = max(worst_eff, d)
            n_eff_equal += d == 0
            worst_p = max(worst_p, float(np.linalg.norm(np.array(xo.pos[:]) - np.array(xb.pos[:]))))
            worst_r = max(worst_r, synth.rot_angle_deg(_rot(xo), _rot(xb)))
    assert n_lio >= scans
    assert worst_p < 1e-3, f"trajectory deviates {worst_p} m"
    assert worst_r < 1e-2, f"attitude deviates {worst_r} deg"
    assert worst_eff <= 3 and n_eff_equal >= 0.98 * n_lio, (worst_eff, n_eff_equal, n_lio)
    mo, mb = o.dump_map(), b.map.dump_map()
    assert_maps_equal(mo, mb, exact=False, rtol=1e-6, what="C1 free-running map")
    return {"case": "C1 free-running", "scans": n_lio, "pts_per_scan": pts, "max_pos_dev_m": worst_p, "max_att_dev_deg": worst_r,
            "scans_with_identical_effect_num": int(n_eff_equal), "max_effect_num_diff": int(worst_eff), "map_voxels": int(len(mo)),
            "seconds": round(time.time() - t0, 1)}


def _teacher_forced(oracle_mod, cfg, pk, min_scans, what, map_check_every=0):
    """Every measurement pass of every scan replayed with the oracle's iteration states (tier 1 + 2), the map fed with the
    oracle's pv_list (tier 1+), for all LIO_MAPPING scans of `pk`."""
    from voxelmapplus_fastlio2_b200.bindings import HotPath
    o = oracle_mod.Oracle(cfg)
    g = HotPath(cfg)
    n_lio = checked = 0
    worst_H = worst_b = worst_res = 0.0
    t0 = time.time()
    for p in pk:
        st = o.lio_process(p.imus, p.cloud, p.t0, p.t1)
        _, _, status = o.lio_state()
        if status == 1:
            continue
        xyz = np.ascontiguousarray(p.cloud[:, :3])
        x0, P0 = o.get_prior()
        if st.iters == 0:
            sg = g.first_scan(x0, P0, xyz)
            assert sg["n_touch"] == st.map.n_touch and sg["n_refit"] == st.map.n_refit and sg["refit_points"] == st.map.refit_points
            pw_o, pc_o = o.dump_world_points()
            pw_g, pc_g = g.dump_world_points()
            assert np.array_equal(pw_o, pw_g) and np.array_equal(pc_o, pc_g), "first scan: world points / covariances differ"
            continue
        g.set_scan(xyz)
        for k in range(st.iters):
            H, b, eff = g.measure(o.get_iter_state(k), P0)
            assert eff == st.effect_num[k], f"scan {p.index} iter {k}: effect_num {eff} vs {st.effect_num[k]}"
            Ho, bo = o.get_iter_Hb(k)
            eH = float(np.abs(H - Ho).max() / np.abs(Ho).max())
            eb = float(np.abs(b - bo).max() / max(np.abs(bo).max(), 1e-300))
            assert eH <= RTOL_T2 and eb <= RTOL_T2, f"scan {p.index} iter {k}: H {eH:.2e} b {eb:.2e}"
            worst_H, worst_b = max(worst_H, eH), max(worst_b, eb)
            checked += 1
        co, cg = o.dump_correspondences(), g.dump_correspondences()
        assert np.array_equal(co["keys"], cg["keys"]), f"scan {p.index}: voxel keys differ"
        assert np.array_equal(co["status"], cg["status"]), f"scan {p.index}: found / is_plane / is_valid differ"
        v = (co["status"] & 4) != 0
        if v.any():
            er = float((np.abs(cg["residual"][v] - co["residual"][v]) / np.maximum(np.abs(co["residual"][v]), 1e-3)).max())
            assert er <= RTOL_T2, f"scan {p.index}: residuals {er:.2e}"
            worst_res = max(worst_res, er)
            assert np.array_equal(co["plane_norm"][v], cg["plane_norm"][v])
        pw, pc = o.dump_world_points()
        sg = g.map_update(pw, pc)
        for f in COUNTERS:
            assert sg[f] == getattr(st.map, f), f"scan {p.index}: {f} {sg[f]} vs {getattr(st.map, f)}"
        assert np.array_equal(o.dump_evicted(), g.dump_evicted()), f"scan {p.index}: evicted keys differ"
        n_lio += 1
        if map_check_every and n_lio % map_check_every == 0:
            assert_maps_equal(o.dump_map(), g.dump_map(), exact=True, what=f"{what}, scan {p.index}")
    assert n_lio >= min_scans, n_lio
    mo = o.dump_map()
    assert_maps_equal(mo, g.dump_map(), exact=True, what=f"{what}, final map")
    return {"case": what, "scans": n_lio, "measurement_passes": checked, "max_rel_err_H": worst_H, "max_rel_err_b": worst_b,
            "max_rel_err_residual": worst_res, "map_voxels": int(len(mo)), "seconds": round(time.time() - t0, 1)}


def c1_teacher_forced(oracle_mod, scans=300, pts=20000):
    """C1 size, teacher-forced: correspondences bit-exact in every pass of every scan, H / b <= 1e-9, map bit-exact."""
    cfg = default_config(max_points_per_scan=pts + 64, voxel_size=0.5, map_capacity=100000)
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=pts))
    return _teacher_forced(oracle_mod, cfg, _packages(seq, scans + 3), scans, "C1 teacher-forced", map_check_every=100)


# ----------------------------------------------------------------------------------------------------------------- C2
def c2_teacher_forced(oracle_mod, scans=50, pts=200000):
    """BASELINE configs[1]: 200 000 pts/scan, 0.25 m voxels, 4 iterations, capacity 400 000; teacher-forced over >= 50 scans (past the
    static start-up of the trajectory: the sensor moves from scan 20 on)."""
    cfg = default_config(max_points_per_scan=pts + 64, voxel_size=0.25, opti_max_iter=4, map_capacity=400000)
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=pts))
    return _teacher_forced(oracle_mod, cfg, _packages(seq, scans + 3), scans, "C2 teacher-forced")


# ----------------------------------------------------------------------------------------------------------------- C3
def c3_city_eviction(oracle_mod, scans=2000, pts=20000, capacity=100000):
    """BASELINE configs[2]: the 5 km city drive of tools/bench_c3.py at its real map capacity; the oracle runs free, the device map
    is fed the oracle's world points scan by scan: counters and the evicted-key stream of every scan and the final map are
    bit-exact (LRU eviction sets in once the map holds `capacity` voxels, around scan 600 at the default sizes)."""
    from voxelmapplus_fastlio2_b200.bindings import HotPath
    cfg = default_config(max_points_per_scan=pts + 64, map_capacity=capacity)
    traj = synth.Trajectory(centre=(600.0, 600.0, 1.8), ax=560.0, ay=2.0, period=448.0)
    seq = synth.Sequence(scene=synth.scene_city(pilasters=True), traj=traj, sensor=synth.SensorConfig(pts_per_scan=pts), seed=0xC3, cull=True)
    pk = _packages(seq, scans + 3)
    o = oracle_mod.Oracle(cfg)
    g = HotPath(cfg)
    evicted = merges = n_lio = 0
    t0 = time.time()
    for p in pk:
        c1 = p.cloud
        so = o.lio_process(p.imus, c1, p.t0, p.t1)
        _, _, s1 = o.lio_state()
        if s1 < 2 and so.map.n_points == 0:
            continue
        pw, pc = o.dump_world_points(len(c1))
        sg = g.map_build(pw, pc) if so.iters == 0 else g.map_update(pw, pc)
        for f in COUNTERS:
            assert sg[f] == getattr(so.map, f), (p.index, f, sg[f], getattr(so.map, f))
        if sg["n_evicted"]:
            assert np.array_equal(o.dump_evicted(), g.dump_evicted()), f"evicted keys differ in scan {p.index}"
        evicted += sg["n_evicted"]
        merges += sg["n_merge"]
        n_lio += 1
    mo = o.dump_map()
    assert_maps_equal(mo, g.dump_map(), exact=True, what="C3 city map, same world points")
    return {"case": "C3 city drive, eviction stream", "scans": n_lio, "pts_per_scan": pts, "map_capacity": capacity, "evicted": int(evicted),
            "merges": int(merges), "map_voxels": int(len(mo)), "seconds": round(time.time() - t0, 1)}


# ----------------------------------------------------------------------------------------------------------------- C4
def c4_map_slice(oracle_mod, points=5_000_000, batch=200_000, capacity=100000):
    """BASELINE configs[3] slice: the map-update microbench's batches (tools/bench_map.py) through VoxelMap::build / update on both
    sides: counters and evicted keys per batch, final map bit-exact."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_map
    from voxelmapplus_fastlio2_b200.bindings import HotPath
    nb = max(3, points // batch)
    batches = bench_map.make_batches(nb, batch)
    cfg = default_config(max_points_per_scan=batch + 64, map_capacity=capacity)
    o = oracle_mod.Oracle(cfg)
    g = HotPath(cfg)
    tot = dict(n_points=0, n_evicted=0, n_merge=0, n_refit=0)
    t0 = time.time()
    for k, (p, c) in enumerate(batches):
        so, sg = (o.map_update(p, c), g.map_update(p, c)) if k else (o.map_build(p, c), g.map_build(p, c))
        assert so == sg, f"batch {k}: counters differ\n{so}\n{sg}"
        if so["n_evicted"]:
            assert np.array_equal(o.dump_evicted(), g.dump_evicted()), f"batch {k}: evicted keys differ"
        for f in tot:
            tot[f] += so[f]
    mo = o.dump_map()
    assert_maps_equal(mo, g.dump_map(), exact=True, what="C4 slice map")
    return {"case": "C4 map-update slice", "batches": nb, "capacity": capacity, **{k: int(v) for k, v in tot.items()}, "map_voxels": int(len(mo)),
            "seconds": round(time.time() - t0, 1)}
