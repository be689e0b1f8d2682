"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tiers (BASELINE.json north_star):
  1  voxel keys, point->plane correspondences (found / is_plane / is_valid), eviction sets,
     map flags / counts / LRU order                                  -> bit-exact
  2  residuals, H, b (fp64)                                          -> 1e-9 relative
  3  trajectory over a run                                           -> 1 mm / 0.01 deg
With teacher forcing (the oracle's point/cov lists handed to vmp_map_update) the plane
parameters themselves are compared bit for bit as well.
"""
import numpy as np
import pytest

from helpers import assert_maps_equal, cov_rel_err, wall_workload
from voxelmapplus_fastlio2_b200 import synth
from voxelmapplus_fastlio2_b200.bindings import HotPath, VmpError
from voxelmapplus_fastlio2_b200.ctypes_defs import default_config
from voxelmapplus_fastlio2_b200.lio import LIOBuilder

pytestmark = pytest.mark.gpu

RTOL_T2 = 1e-9      # tier 2 tolerance (relative), stated by north_star


def _pair(oracle_mod, **kw):
    cfg = default_config(**kw)
    return oracle_mod.Oracle(cfg), HotPath(cfg)


# --------------------------------------------------------------------------- map update
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_map_update_bitexact_merge(oracle_mod, seed):
    """fill -> refit cadence -> close -> merge (Q7-Q13), planes compared bit for bit."""
    o, g = _pair(oracle_mod, max_point_thresh=30, update_size_thresh=5, map_capacity=100000, max_points_per_scan=4096)
    work = wall_workload(seed)
    merges = 0
    for s, (p, c) in enumerate(work):
        if s == 0:
            so, sg = o.map_build(p, c), g.map_build(p, c)
        else:
            so, sg = o.map_update(p, c), g.map_update(p, c)
        assert so == sg, f"scan {s}: counters differ\n{so}\n{sg}"
        merges += so["n_merge"]
        assert_maps_equal(o.dump_map(), g.dump_map(), exact=True, what=f"seed {seed} scan {s}")
    assert merges > 0, "workload did not exercise merge()"


@pytest.mark.parametrize("voxel_size", [0.3, 0.7])
def test_voxel_sizes_that_are_not_powers_of_two(oracle_mod, voxel_size):
    """The voxel key is floor(p / voxel_size) (voxel_map.cpp:194-198).  For the power-of-two sizes of the BASELINE configurations the
    device multiplies by the exact reciprocal (the same value bit for bit); any other size takes the division.  Keys, correspondences and
    maps stay bit-exact on that path too: map update on the oracle's world points, and one measurement pass per scan."""
    o, g = _pair(oracle_mod, voxel_size=voxel_size, max_points_per_scan=4096)
    for s, (p, c) in enumerate(wall_workload(21, scans=14, pts=3000)):
        so, sg = (o.map_update(p, c), g.map_update(p, c)) if s else (o.map_build(p, c), g.map_build(p, c))
        assert so == sg, f"scan {s}: counters differ\n{so}\n{sg}"
    assert_maps_equal(o.dump_map(), g.dump_map(), exact=True, what=f"voxel_size {voxel_size}")
    # points on and next to voxel faces: the keys of the measurement pass
    rng = np.random.default_rng(5)
    k = rng.integers(-12, 12, (3000, 3)).astype(np.float64)
    pts = (k * voxel_size + rng.choice([0.0, 1e-7, -1e-7, 0.4 * voxel_size], (3000, 3))).astype(np.float32)
    from voxelmapplus_fastlio2_b200.ctypes_defs import VmpState
    x0 = VmpState.identity()
    P0 = np.eye(23) * 1e-4
    for h in (o, g):
        h.set_scan(pts)
        h.measure(x0, P0)
    co, cg = o.dump_correspondences(len(pts)), g.dump_correspondences(len(pts))
    assert np.array_equal(co["keys"], cg["keys"]) and np.array_equal(co["status"], cg["status"])
    assert (co["status"] & 1).any()                       # some of the probes hit voxels of the map


def test_map_update_default_thresholds(oracle_mod):
    o, g = _pair(oracle_mod, max_points_per_scan=4096)
    for s, (p, c) in enumerate(wall_workload(11, scans=25, pts=3000)):
        so, sg = (o.map_update(p, c), g.map_update(p, c)) if s else (o.map_build(p, c), g.map_build(p, c))
        assert so == sg, f"scan {s}: counters differ\n{so}\n{sg}"
    assert_maps_equal(o.dump_map(), g.dump_map(), exact=True, what="default thresholds")


def _moving_workload(seed, scans, pts):
    rng = np.random.Generator(np.random.Philox(key=seed))
    out = []
    for s in range(scans):
        # a corridor that advances 1 m per scan: old voxels fall out of view and get evicted
        x0 = 1.0 * s
        p = np.concatenate([
            np.stack([rng.uniform(x0, x0 + 6, pts // 2), rng.normal(0.25, 0.01, pts // 2), rng.uniform(0, 2, pts // 2)], 1),
            np.stack([rng.uniform(x0, x0 + 6, pts // 2), rng.uniform(0, 3, pts // 2), rng.normal(0.1, 0.01, pts // 2)], 1)])
        p = p[rng.permutation(len(p))].astype(np.float32).astype(np.float64)
        c = np.tile((np.eye(3) * 1e-4).reshape(1, 9), (len(p), 1))
        out.append((p, c))
    return out


def test_map_update_lru_eviction(oracle_mod):
    """capacity smaller than the trajectory's voxel count: eviction order must be the reference's (Q17)."""
    o, g = _pair(oracle_mod, max_point_thresh=20, update_size_thresh=5, map_capacity=400, max_points_per_scan=4096)
    evicted_total = 0
    for s, (p, c) in enumerate(_moving_workload(5, 40, 1200)):
        so, sg = (o.map_update(p, c), g.map_update(p, c)) if s else (o.map_build(p, c), g.map_build(p, c))
        assert so == sg, f"scan {s}: counters differ\n{so}\n{sg}"
        eo, eg = o.dump_evicted(), g.dump_evicted()
        assert np.array_equal(eo, eg), f"scan {s}: eviction order differs"
        evicted_total += len(eo)
        assert_maps_equal(o.dump_map(), g.dump_map(), exact=True, what=f"eviction scan {s}")
    assert evicted_total > 100


def test_map_update_evict_and_recreate_same_scan(oracle_mod):
    """a voxel that is the LRU victim early in a scan and is hit again later in the same scan."""
    o, g = _pair(oracle_mod, max_point_thresh=20, update_size_thresh=5, map_capacity=60, max_points_per_scan=4096)
    rng = np.random.Generator(np.random.Philox(key=9))
    cov = lambda n: np.tile((np.eye(3) * 1e-4).reshape(1, 9), (n, 1))
    # scan 0: 60 distinct voxels along x (fills the map exactly), oldest = voxel 0
    p0 = np.stack([np.arange(60) * 0.5 + 0.25, np.full(60, 0.25), np.full(60, 0.25)], 1)
    # scan 1: 10 new voxels first (evicting voxels 0..9), then points into voxels 3 and 5 again, then old ones
    new = np.stack([np.arange(100, 110) * 0.5 + 0.25, np.full(10, 0.25), np.full(10, 0.25)], 1)
    again = np.stack([np.array([3, 5, 3, 20, 21]) * 0.5 + 0.3, np.full(5, 0.3), np.full(5, 0.2)], 1)
    p1 = np.concatenate([new, again, new + 0.01])
    p3 = np.stack([rng.integers(200, 230, 90) * 0.5 + 0.25, np.full(90, 0.25), rng.integers(0, 2, 90) * 0.5 + 0.25], 1)
    for s, p in enumerate([p0, p1, p0[::-1].copy(), p3, p1, p0]):
        p = p.astype(np.float32).astype(np.float64)
        so, sg = o.map_update(p, cov(len(p))), g.map_update(p, cov(len(p)))
        assert so == sg, f"scan {s}: counters differ\n{so}\n{sg}"
        assert np.array_equal(o.dump_evicted(), g.dump_evicted()), f"scan {s}: eviction order differs"
        assert_maps_equal(o.dump_map(), g.dump_map(), exact=True, what=f"recreate scan {s}")


def test_many_recreations_in_one_scan(oracle_mod):
    """150 voxels are LRU victims early in a scan and are all hit again later in the same scan (a drive that turns back into
    the oldest part of its map: the C3 city run overflowed a 64-entry queue here).  Every re-creation evicts a further voxel."""
    o, g = _pair(oracle_mod, max_point_thresh=20, update_size_thresh=5, map_capacity=400, max_points_per_scan=4096)
    cov = lambda n: np.tile((np.eye(3) * 1e-4).reshape(1, 9), (n, 1))
    line = lambda idx, y=0.25: np.stack([np.asarray(idx) * 0.5 + 0.25, np.full(len(idx), y), np.full(len(idx), 0.25)], 1)
    p0 = line(np.arange(400))                                     # fills the map exactly; oldest = voxel 0
    p1 = np.concatenate([line(np.arange(1000, 1150)),             # 150 creations evict voxels 0..149 ...
                         line(np.arange(150), 0.3),               # ... which are then hit again (re-created, evicting 150..299)
                         line(np.arange(1000, 1150), 0.2)])
    p2 = np.concatenate([line(np.arange(2000, 2100)), line(np.arange(300, 400), 0.3), line(np.arange(0, 150, 2), 0.35)])
    for s, p in enumerate([p0, p1, p2, p0]):
        p = p.astype(np.float32).astype(np.float64)
        so, sg = o.map_update(p, cov(len(p))), g.map_update(p, cov(len(p)))
        assert so == sg, f"scan {s}: counters differ\n{so}\n{sg}"
        assert np.array_equal(o.dump_evicted(), g.dump_evicted()), f"scan {s}: eviction order differs"
        assert_maps_equal(o.dump_map(), g.dump_map(), exact=True, what=f"many re-creations, scan {s}")
        if s == 1:
            assert so["n_evicted"] == 300 and so["n_created"] == 300


def test_merge_serial_redo_is_exact(oracle_mod, monkeypatch):
    """The exact fallback of the merge phase (vmp_merge.cuh): with VMP_MERGE_MAX_DEPTH=-1 every scan in which a merge succeeds takes the
    parallel rounds back from the undo log and redoes its merge() calls one at a time in strict event order.  Same bit-exact map."""
    monkeypatch.setenv("VMP_MERGE_MAX_DEPTH", "-1")
    o, g = _pair(oracle_mod, max_point_thresh=30, update_size_thresh=5, map_capacity=100000, max_points_per_scan=4096)
    monkeypatch.delenv("VMP_MERGE_MAX_DEPTH")
    merges = redone = 0
    for s, (p, c) in enumerate(wall_workload(2, scans=14)):
        so, sg = (o.map_update(p, c), g.map_update(p, c)) if s else (o.map_build(p, c), g.map_build(p, c))
        assert so == sg, f"scan {s}: counters differ\n{so}\n{sg}"
        merges += so["n_merge"]
        redone += (g.debug_counters()[2] >> 30) & 1
        assert_maps_equal(o.dump_map(), g.dump_map(), exact=True, what=f"serial redo, scan {s}")
    assert merges > 0 and redone > 0, (merges, redone)


def test_merge_active_set_beyond_the_shared_arrays(oracle_mod, monkeypatch):
    """~7700 coplanar voxels close and start merging within the same two scans: more than a thousand voxels are active in the merge phase at
    once.  With the shared arrays of the parallel rounds limited to 256 entries (VMP_MERGE_CAP; 2048 in production) such scans run in the
    exact serial mode from the start, and scans whose set outgrows the arrays on the way are taken back and redone (round 1: E_MERGE_CAP)."""
    monkeypatch.setenv("VMP_MERGE_CAP", "256")
    o, g = _pair(oracle_mod, max_point_thresh=20, update_size_thresh=5, map_capacity=100000, max_points_per_scan=65536)
    monkeypatch.delenv("VMP_MERGE_CAP")
    rng = np.random.Generator(np.random.Philox(key=77))
    merges = 0
    peak_active = 0
    for s in range(4):
        n = 60000
        p = np.stack([rng.uniform(0, 48, n), rng.uniform(0, 40, n), rng.normal(0.12, 0.004, n)], 1).astype(np.float32).astype(np.float64)
        c = np.tile((np.eye(3) * 1e-4).reshape(1, 9), (n, 1))
        so, sg = (o.map_update(p, c), g.map_update(p, c)) if s else (o.map_build(p, c), g.map_build(p, c))
        assert so == sg, f"scan {s}: counters differ\n{so}\n{sg}"
        merges += so["n_merge"]
        peak_active = max(peak_active, g.debug_counters()[0])
        assert_maps_equal(o.dump_map(), g.dump_map(), exact=True, what=f"large active set, scan {s}")
    assert merges > 1000 and peak_active > 1000, (merges, peak_active)


def test_unkeyable_points_are_counted_skips(oracle_mod):
    """NaN / inf / out-of-range (> 2^20 voxels) world points are skipped and counted, they neither corrupt voxel (0,0,0) nor leave a
    sticky error behind: the device map fed scan + 4 bad points equals the oracle map fed the scan alone, later updates return OK."""
    o, g = _pair(oracle_mod, max_point_thresh=30, update_size_thresh=5, map_capacity=100000, max_points_per_scan=4096)
    bad = np.array([[np.nan, 0.1, 0.1], [np.inf, 1.0, 1.0], [1.0e9, 0.0, 0.0], [0.2, -np.inf, np.nan]])
    cb = np.tile((np.eye(3) * 1e-4).reshape(1, 9), (len(bad), 1))
    for s, (p, c) in enumerate(wall_workload(21, scans=6, pts=1500)):
        so = o.map_update(p, c) if s else o.map_build(p, c)
        if s in (1, 3):
            pg, cg = np.concatenate([p, bad]), np.concatenate([c, cb])        # appended: the point indices of the real points are unchanged
        else:
            pg, cg = p, c
        sg = g.map_update(pg, cg) if s else g.map_build(pg, cg)
        assert sg["n_skipped"] == (len(bad) if s in (1, 3) else 0)
        for f in ("n_ins", "n_touch", "n_created", "n_refit", "refit_points", "n_full", "n_mergeprobe", "n_merge", "n_evicted", "map_size"):
            assert sg[f] == so[f], (s, f)
        assert_maps_equal(o.dump_map(), g.dump_map(), exact=True, what=f"unkeyable points, scan {s}")


def test_capacity_smaller_than_scan_fails_loudly(oracle_mod):
    """documented restriction: the LRU victim must not have been touched in the same scan."""
    cfg = default_config(map_capacity=8, max_points_per_scan=1024)
    g = HotPath(cfg)
    p = np.stack([np.arange(64) * 0.5 + 0.25, np.zeros(64), np.zeros(64)], 1)
    p = np.concatenate([p, p])
    with pytest.raises(VmpError, match="map_capacity"):
        g.map_update(p, np.tile(np.eye(3).reshape(1, 9) * 1e-4, (len(p), 1)))
    # error bits are per update: the condition is reported once, a later update that fits starts clean
    q = np.stack([np.arange(4) * 0.5 + 0.25, np.zeros(4), np.zeros(4)], 1)
    g.map_update(q, np.tile(np.eye(3).reshape(1, 9) * 1e-4, (len(q), 1)))


# --------------------------------------------------------------------------- measurement model + whole scan
def _sequence(pts=3000, scans=14):
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=pts))
    return list(seq.packages(scans))


@pytest.mark.parametrize("estimate_ext", [0, 1])
def test_measure_and_scan_teacher_forced(oracle_mod, estimate_ext):
    """Every measurement pass of every scan replayed with the oracle's states: correspondences bit-exact,
    H / b to 1e-9; map fed with the oracle's pv_list: bit-exact planes; vmp_scan posterior vs oracle."""
    kw = dict(max_points_per_scan=4096, estimate_ext=estimate_ext)
    cfg = default_config(**kw)
    o = oracle_mod.Oracle(cfg)
    o.track_margins(True)
    g_tf = HotPath(cfg)          # teacher-forced: measure + map_update with the oracle's data
    g_fr = HotPath(cfg)          # vmp_scan with the oracle's prior, its own posterior / map
    checked_iters = 0
    for pk in _sequence():
        st = o.lio_process(pk.imus, pk.cloud, pk.t0, pk.t1)
        x_post, P_post, status = o.lio_state()
        if status == 1:
            continue
        xyz = np.ascontiguousarray(pk.cloud[:, :3])          # undistorted in place by the oracle
        x0, P0 = o.get_prior()
        if st.iters == 0:                                    # MAP_INIT scan
            so = st.map
            sg = g_tf.first_scan(x0, P0, xyz)
            g_fr.first_scan(x0, P0, xyz)
            assert sg["n_touch"] == so.n_touch and sg["n_refit"] == so.n_refit
            pw_o, pc_o = o.dump_world_points()
            pw_g, pc_g = g_tf.dump_world_points()
            assert np.array_equal(pw_o, pw_g), "float32 world points differ"
            assert np.array_equal(pc_o, pc_g), "world covariances differ"
            assert_maps_equal(o.dump_map(), g_tf.dump_map(), exact=True, what="first scan")
            continue
        # --- teacher-forced measurement passes
        g_tf.set_scan(xyz)
        for k in range(st.iters):
            xk = o.get_iter_state(k)
            H, b, eff = g_tf.measure(xk, P0)
            assert eff == st.effect_num[k], f"scan {pk.index} iter {k}: effect_num {eff} vs {st.effect_num[k]}"
            Ho, bo = o.get_iter_Hb(k)            # what the oracle's sharedUpdateFunc returned for this pass
            scale = np.abs(Ho).max()
            assert np.abs(H - Ho).max() <= RTOL_T2 * scale, f"H differs: {np.abs(H - Ho).max() / scale:.3e}"
            assert np.abs(b - bo).max() <= RTOL_T2 * max(np.abs(bo).max(), 1e-300), "b differs"
            checked_iters += 1
        # records after the last pass must agree with the oracle's records
        co, cg = o.dump_correspondences(), g_tf.dump_correspondences()
        assert np.array_equal(co["keys"], cg["keys"]), "voxel keys differ"
        assert np.array_equal(co["status"], cg["status"]), "found/is_plane/is_valid differ"
        v = (co["status"] & 4) != 0
        np.testing.assert_allclose(cg["residual"][v], co["residual"][v], rtol=RTOL_T2, atol=1e-12)
        assert np.array_equal(co["plane_norm"][v], cg["plane_norm"][v])
        # --- teacher-forced map update
        pw, pc = o.dump_world_points()
        sg = g_tf.map_update(pw, pc)
        for f in ("n_ins", "n_touch", "n_created", "n_refit", "refit_points", "n_full", "n_mergeprobe", "n_merge", "n_evicted", "map_size"):
            assert sg[f] == getattr(st.map, f), f"scan {pk.index}: {f} {sg[f]} vs {getattr(st.map, f)}"
        assert_maps_equal(o.dump_map(), g_tf.dump_map(), exact=True, what=f"scan {pk.index}")
        # --- vmp_scan from the oracle's prior
        xg, Pg, sgs = g_fr.scan(x0, P0, xyz)
        assert sgs.iters == st.iters and list(sgs.effect_num[:st.iters]) == list(st.effect_num[:st.iters])
        assert np.abs(np.array(xg.pos[:]) - np.array(x_post.pos[:])).max() < 1e-9
        assert np.abs(np.array(xg.rot[:]) - np.array(x_post.rot[:])).max() < 1e-10
        # posterior covariance: the device evaluates IESKF::update through the matrix-inversion lemma (vmp_solve.cuh), an algebraically
        # identical form; on identical inputs both are within ~1e-13 of the 50-digit value (tests/test_posterior_precision.py).  g_fr keeps
        # its OWN map (float32 world points of its own posteriors), so its H differs from the oracle's in the 8th digit in these first
        # scans (P = identity at the start, five iterations): measured <= 1e-8 here, 1e-7 asserted (round 1: rtol 2e-5)
        assert cov_rel_err(Pg, P_post) <= 1e-7, cov_rel_err(Pg, P_post)
    assert checked_iters > 15
    # SURVEY.md 8(c) safeguard (iii): none of the bit-exact decisions above was a tie of the arithmetic
    assert min(o.gate_margins().values()) > 1e-10, o.gate_margins()


def test_dense_scan_c2_size(oracle_mod):
    """BASELINE.json configs[1] size: 200 000 pts/scan, 0.25 m voxels, 4 iterations — every tier at full size for two scans
    (heavy voxels with > 1000 points per scan exercise the in-place bitonic sort and the build overflow path, Q18)."""
    n = 200000
    cfg = default_config(max_points_per_scan=n + 64, voxel_size=0.25, opti_max_iter=4, map_capacity=400000)
    o = oracle_mod.Oracle(cfg)
    g = HotPath(cfg)
    g2 = HotPath(cfg)                 # map only, fed the oracle's own world points: bit-exact at full size
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=n))
    scans = 0
    for pk in seq.packages(4):
        st = o.lio_process(pk.imus, pk.cloud, pk.t0, pk.t1)
        x_post, P_post, status = o.lio_state()
        if status == 1:
            continue
        xyz = np.ascontiguousarray(pk.cloud[:, :3])
        x0, P0 = o.get_prior()
        pw, pc = o.dump_world_points(len(xyz))
        s2 = g2.map_build(pw, pc) if st.iters == 0 else g2.map_update(pw, pc)
        for f in ("n_ins", "n_touch", "n_created", "n_refit", "refit_points", "n_full", "n_mergeprobe", "n_merge", "map_size"):
            assert s2[f] == getattr(st.map, f), f
        if st.iters == 0:
            sg = g.first_scan(x0, P0, xyz)
            assert sg["n_touch"] == st.map.n_touch and sg["n_refit"] == st.map.n_refit and sg["refit_points"] == st.map.refit_points
            continue
        xg, Pg, sgs = g.scan(x0, P0, xyz)
        assert sgs.iters == st.iters and list(sgs.effect_num[:st.iters]) == list(st.effect_num[:st.iters])
        co, cg = o.dump_correspondences(), g.dump_correspondences()
        assert np.array_equal(co["keys"], cg["keys"]) and np.array_equal(co["status"], cg["status"])
        # 200 000 points make H ~ 1e9: the posterior agrees with the oracle to ~1e-9 m; 1e-7 m (0.1 um) is asserted
        assert np.abs(np.array(xg.pos[:]) - np.array(x_post.pos[:])).max() < 1e-7
        for f in ("n_ins", "n_touch", "n_created", "n_refit", "refit_points", "n_full", "n_mergeprobe", "n_merge", "map_size"):
            assert getattr(sgs.map, f) == getattr(st.map, f), f
        scans += 1
    assert scans == 2
    assert_maps_equal(o.dump_map(), g2.dump_map(), exact=True, what="dense map, same world points")
    assert_maps_equal(o.dump_map(), g.dump_map(), exact=False, what="dense scan map")


def test_lio_trajectory(oracle_mod):
    """Tier 3: free-running estimators (host predict/undistort + device update) stay within 1 mm / 0.01 deg."""
    cfg = default_config(max_points_per_scan=8192)
    o = oracle_mod.Oracle(cfg)
    b = LIOBuilder(cfg)
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=6000))
    worst_p, worst_r = 0.0, 0.0
    for pk in seq.packages(60):
        c1, c2 = pk.cloud.copy(), pk.cloud.copy()
        so = o.lio_process(pk.imus, c1, pk.t0, pk.t1)
        sb = b.process(pk.imus, c2, pk.t0, pk.t1)
        xo, Po, s1 = o.lio_state()
        xb, Pb, s2 = b.state()
        assert s1 == s2
        # undistortion is float32 arithmetic on the (1e-9-close) posteriors: identical up to single float32 ulps
        assert np.allclose(c1, c2, rtol=0, atol=2 * float(np.finfo(np.float32).eps) * float(np.abs(c1[:, :3]).max())), "undistorted clouds differ"
        assert (c1 != c2).mean() < 0.1          # flip rate ~ posterior difference / float32 ulp ~ 1e-9 / 6e-8
        if s1 == 2 and so.iters:
            assert sb.iters == so.iters
            worst_p = max(worst_p, float(np.linalg.norm(np.array(xo.pos[:]) - np.array(xb.pos[:]))))
            worst_r = max(worst_r, synth.rot_angle_deg(np.array(xo.rot[:]).reshape(3, 3), np.array(xb.rot[:]).reshape(3, 3)))
    assert worst_p < 1e-3, f"trajectory deviates {worst_p} m"
    assert worst_r < 1e-2, f"attitude deviates {worst_r} deg"
    assert_maps_equal(o.dump_map(), b.map.dump_map(), exact=False, rtol=1e-6, what="free-running map")


def test_lio_trajectory_against_the_python_implementation():
    """Tier 3 without the C++ oracle in the loop: the CUDA path (device compensation + update, host IMU propagation) free-running beside
    tests/lio_pyref.py, the independent Python / numpy / LAPACK implementation of LIOBuilder::process that pins the oracle on the CPU
    (tests/test_oracle_independent.py).  Same iteration counts, pose within 1 mm / 0.01 deg, and the measured deviation printed."""
    from lio_pyref import LioPy
    cfg = default_config(max_points_per_scan=2048, map_capacity=100000)
    b = LIOBuilder(cfg)
    py = LioPy(cfg)
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=700))
    worst_p = worst_r = 0.0
    updates = same_effect = 0
    for pk in seq.packages(32):
        sb = b.process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        py.process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        xb, Pb, s2 = b.state()
        assert s2 == py.status
        if s2 == 2 and sb.iters:
            assert sb.iters == py.iters
            same_effect += list(sb.effect_num[:sb.iters]) == py.effect
            worst_p = max(worst_p, float(np.linalg.norm(np.array(xb.pos[:]) - py.x["pos"])))
            worst_r = max(worst_r, synth.rot_angle_deg(np.array(xb.rot[:]).reshape(3, 3), py.x["rot"]))
            updates += 1
    print(f"CUDA path vs python LIO over {updates} updates: position {worst_p:.3e} m, attitude {worst_r:.3e} deg, effect_num equal in {same_effect}")
    assert updates >= 20 and same_effect >= updates - 2
    assert worst_p < 1e-3 and worst_r < 1e-2, (worst_p, worst_r)


def test_pipelined_mode_is_identical():
    """vmp_set_pipelined: the same posteriors and the same map bit for bit; map counters arrive one scan late."""
    cfg = default_config(max_points_per_scan=8192)
    a = LIOBuilder(cfg)
    b = LIOBuilder(cfg, pipelined=True)
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=6000))
    prev = None
    for pk in seq.packages(40):
        sa = a.process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        sb = b.process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        xa, Pa, s1 = a.state()
        xb, Pb, s2 = b.state()
        assert s1 == s2 and bytes(xa) == bytes(xb) and np.array_equal(Pa, Pb)
        assert sa.iters == sb.iters and list(sa.effect_num) == list(sb.effect_num)
        if sa.iters:
            if prev is not None and prev.iters:
                assert sb.map.as_dict() == prev.map.as_dict()      # lagging by exactly one scan
            prev = sa
    b.map.sync()
    assert_maps_equal(a.map.dump_map(), b.map.dump_map(), exact=True, what="pipelined map")


@pytest.mark.parametrize("estimate_ext", [False, True])
def test_resident_iteration_loop_is_identical(monkeypatch, estimate_ext):
    """The IEKF iterations of a scan as ONE resident launch (k_iekf_loop: solver CTA and measurement CTAs hand the state over through
    a release / acquire flag) against one launch per iteration (VMP_IEKF_LOOP=0): same arithmetic in the same order, so the
    posteriors, the iteration counts, effect_num and the map are identical bit for bit (ieskf.cpp:125-156)."""
    cfg = default_config(max_points_per_scan=8192, estimate_ext=1 if estimate_ext else 0)
    a = LIOBuilder(cfg)
    monkeypatch.setenv("VMP_IEKF_LOOP", "0")
    b = LIOBuilder(cfg)
    monkeypatch.delenv("VMP_IEKF_LOOP")
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=6000))
    iters = []
    for pk in seq.packages(40):
        sa = a.process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        sb = b.process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        xa, Pa, s1 = a.state()
        xb, Pb, s2 = b.state()
        assert s1 == s2 and bytes(xa) == bytes(xb) and np.array_equal(Pa, Pb)
        assert sa.iters == sb.iters and list(sa.effect_num) == list(sb.effect_num) and sa.converged == sb.converged
        assert sa.map.as_dict() == sb.map.as_dict()
        iters.append(sa.iters)
    assert len(set(i for i in iters if i)) >= 2          # scans that stop early and scans that do not
    assert_maps_equal(a.map.dump_map(), b.map.dump_map(), exact=True, what="resident loop vs one launch per iteration")


@pytest.mark.parametrize("chunk_min,pipelined,loop", [(32768, False, "1"), (4096, False, "1"), (4096, True, "1"), (4096, False, "0")])
def test_streamed_upload_is_identical(monkeypatch, chunk_min, pipelined, loop):
    """Streamed upload (VMP_UPLOAD_GATE=1; an experiment that is off by default): vmp_scan with a pageable pointer launches the
    graph behind the FIRST chunk of points; the other chunks travel on their own stream while the first measurement pass runs and
    release its waiting warps one by one (DevCtl::up_pub, written by a stream memory operation behind each chunk).  Only loads are
    delayed: posterior, counters and map are those of the plain upload (the graph waits for all of it), bit for bit.
    VMP_UPLOAD_CHUNK_MIN makes small scans travel in chunks."""
    cfg = default_config(max_points_per_scan=8192)
    monkeypatch.setenv("VMP_UPLOAD_CHUNK_MIN", str(chunk_min))
    monkeypatch.setenv("VMP_IEKF_LOOP", loop)
    monkeypatch.setenv("VMP_UPLOAD_GATE", "1")
    a = HotPath(cfg)
    monkeypatch.delenv("VMP_UPLOAD_GATE")
    b = HotPath(cfg)
    monkeypatch.delenv("VMP_UPLOAD_CHUNK_MIN"); monkeypatch.delenv("VMP_IEKF_LOOP")
    if pipelined:
        a.set_pipelined(True)
    lio = LIOBuilder(cfg, device_undistort=False)          # produces priors and compensated clouds
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=6000))
    n = 0
    for pk in seq.packages(30):
        cloud = pk.cloud.copy()
        st = lio.process(pk.imus, cloud, pk.t0, pk.t1)
        _, _, status = lio.state()
        if status < 2:
            continue
        x0, P0 = lio.prior()
        xyz = np.ascontiguousarray(cloud[:, :3])
        if st.iters == 0:
            a.first_scan(x0, P0, xyz); b.first_scan(x0, P0, xyz)
            continue
        xa, Pa, sa = a.scan(x0, P0, xyz)
        xb, Pb, sb = b.scan(x0, P0, xyz)
        assert bytes(xa) == bytes(xb) and np.array_equal(Pa, Pb) and sa.iters == sb.iters
        assert list(sa.effect_num) == list(sb.effect_num)
        if not pipelined:
            assert sa.map.as_dict() == sb.map.as_dict()
        n += 1
    assert n >= 20
    if pipelined:
        a.sync()
    assert_maps_equal(a.dump_map(), b.dump_map(), exact=True, what="streamed upload")


def test_scan_buffer_fill_is_identical():
    """vmp_scan_buffer_fill (x y z out of the caller's 4- or 12-float point records, by the staging helpers) + vmp_scan_staged == vmp_scan on the
    packed x y z, bit for bit; a stride below 3 is rejected."""
    cfg = default_config(max_points_per_scan=8192)
    a, b = HotPath(cfg), HotPath(cfg)
    lio = LIOBuilder(cfg, device_undistort=False)          # produces priors and compensated clouds
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=5000))
    n = 0
    for pk in seq.packages(10):
        cloud = pk.cloud.copy()
        st = lio.process(pk.imus, cloud, pk.t0, pk.t1)
        _, _, status = lio.state()
        if status < 2:
            continue
        x0, P0 = lio.prior()
        xyz = np.ascontiguousarray(cloud[:, :3])
        if st.iters == 0:
            a.first_scan(x0, P0, xyz); b.first_scan(x0, P0, xyz)
            continue
        rec = cloud if n % 2 == 0 else np.concatenate([cloud[:, :3], np.full((len(cloud), 9), 7.0, np.float32)], 1)     # x y z t  /  a 48-byte PCL-style record
        xa, Pa, sa = a.scan(x0, P0, xyz)
        xb, Pb, sb = b.scan_filled(x0, P0, rec)
        assert bytes(xa) == bytes(xb) and np.array_equal(Pa, Pb) and sa.iters == sb.iters
        assert sa.map.as_dict() == sb.map.as_dict()
        n += 1
    assert n >= 5
    assert_maps_equal(a.dump_map(), b.dump_map(), exact=True, what="filled scan")
    with pytest.raises(VmpError):
        b.scan_filled(x0, P0, cloud, stride=2)


def test_staged_scan_is_identical():
    """vmp_scan_buffer + vmp_scan_staged (points written straight into the pinned staging) == vmp_scan, bit for bit."""
    cfg = default_config(max_points_per_scan=8192)
    a, b = HotPath(cfg), HotPath(cfg)
    lio = LIOBuilder(cfg, device_undistort=False)          # produces priors and compensated clouds
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=5000))
    n = 0
    for pk in seq.packages(10):
        cloud = pk.cloud.copy()
        st = lio.process(pk.imus, cloud, pk.t0, pk.t1)
        _, _, status = lio.state()
        if status < 2:
            continue
        x0, P0 = lio.prior()
        xyz = np.ascontiguousarray(cloud[:, :3])
        if st.iters == 0:
            a.first_scan(x0, P0, xyz); b.first_scan(x0, P0, xyz)
            continue
        xa, Pa, sa = a.scan(x0, P0, xyz)
        xb, Pb, sb = b.scan_staged(x0, P0, xyz)
        assert bytes(xa) == bytes(xb) and np.array_equal(Pa, Pb) and sa.iters == sb.iters
        assert sa.map.as_dict() == sb.map.as_dict()
        # ... and the host LIOBuilder, which stages its scans the same way, got the same posterior
        xl, Pl, _ = lio.state()
        assert bytes(xl) == bytes(xa) and np.array_equal(Pl, Pa)
        n += 1
    assert n >= 5
    assert_maps_equal(a.dump_map(), b.dump_map(), exact=True, what="staged vs copied scan")


def test_pipelined_reports_capacity_error_late():
    """a map update that exhausts the capacity while pipelined surfaces on the next call on the handle"""
    cfg = default_config(max_points_per_scan=4096, map_capacity=64)
    g = HotPath(cfg)
    rng = np.random.default_rng(5)
    from voxelmapplus_fastlio2_b200.ctypes_defs import VmpState
    x = VmpState(); x.rot[0] = x.rot[4] = x.rot[8] = 1.0; x.rot_ext[0] = x.rot_ext[4] = x.rot_ext[8] = 1.0; x.g[2] = -9.81
    P = np.eye(23) * 1e-4
    pts = rng.uniform(-4, 4, (32, 3)).astype(np.float32)
    g.first_scan(x, P, pts)
    g.set_pipelined(True)
    big = rng.uniform(-30, 30, (4000, 3)).astype(np.float32)          # touches far more than 64 voxels in one scan
    g.scan(x, P, big)                                                 # posterior delivered, the update fails behind it
    with pytest.raises(VmpError):
        g.sync()


def test_device_undistortion_matches_host(oracle_mod):
    """the point loop of undistortCloud on the device (vmp_scan_raw) against the host loop and the oracle;
    the second half of the run feeds time-shuffled clouds (the sort of lio_builder.cpp:75)."""
    cfg = default_config(max_points_per_scan=8192)
    o = oracle_mod.Oracle(cfg)
    dev = LIOBuilder(cfg)                                # device compensation (default)
    host = LIOBuilder(cfg, device_undistort=False)       # host compensation, as before
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=5000))
    rng = np.random.default_rng(11)
    eps = float(np.finfo(np.float32).eps)
    for pk in seq.packages(30):
        cloud = pk.cloud.copy()
        if pk.index >= 15:
            cloud = np.ascontiguousarray(cloud[rng.permutation(len(cloud))])
        c0, c1, c2 = cloud.copy(), cloud.copy(), cloud.copy()
        so = o.lio_process(pk.imus, c0, pk.t0, pk.t1)
        sd = dev.process(pk.imus, c1, pk.t0, pk.t1)
        sh = host.process(pk.imus, c2, pk.t0, pk.t1)
        xo, _, s0 = o.lio_state()
        xd, _, s1 = dev.state()
        xh, _, s2 = host.state()
        assert s0 == s1 == s2
        if s0 < 2 or so.iters == 0:
            continue
        assert sd.iters == so.iters == sh.iters
        tol = 2 * eps * float(np.abs(c0[:, :3]).max())
        assert np.array_equal(c0[:, 3], c1[:, 3]) and np.array_equal(c0[:, 3], c2[:, 3])      # same time order
        assert np.allclose(c1, c0, rtol=0, atol=tol) and np.allclose(c1, c2, rtol=0, atol=tol)
        assert np.linalg.norm(np.array(xd.pos[:]) - np.array(xo.pos[:])) < 1e-3
        assert np.linalg.norm(np.array(xd.pos[:]) - np.array(xh.pos[:])) < 1e-3
    assert_maps_equal(o.dump_map(), dev.map.dump_map(), exact=False, rtol=1e-6, what="device-undistorted map")


def test_device_imu_propagation_matches_host_and_oracle(oracle_mod):
    """IESKF::predict + the IMU pose list on the device (vmp_scan_raw_predict, k_predict), state / P resident there.
    Free-running against the host-propagating builder and the oracle: the propagated prior of every scan against the oracle's
    (IESKF::predict, ieskf.cpp:101-123), the posteriors within tier 3, identical iteration counts."""
    from helpers import cov_rel_err
    cfg = default_config(max_points_per_scan=8192)
    o = oracle_mod.Oracle(cfg)
    dev = LIOBuilder(cfg, device_predict=True)
    host = LIOBuilder(cfg)
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=5000))
    worst_x = worst_P = worst_post = 0.0
    n = 0
    for pk in seq.packages(45):
        so = o.lio_process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        sd = dev.process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        sh = host.process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        xo, _, s0 = o.lio_state()
        xd, Pd, s1 = dev.state()
        xh, Ph, s2 = host.state()
        assert s0 == s1 == s2
        if s0 < 2 or so.iters == 0:
            continue
        assert sd.iters == so.iters == sh.iters
        xp_o, Pp_o = o.get_prior()
        xp_d, Pp_d = dev.map.get_prior()
        for f in ("pos", "rot", "vel", "bg", "ba", "g", "rot_ext", "pos_ext"):
            worst_x = max(worst_x, float(np.abs(np.array(getattr(xp_d, f)[:]) - np.array(getattr(xp_o, f)[:])).max()))
        worst_P = max(worst_P, cov_rel_err(Pp_d, Pp_o))
        worst_post = max(worst_post, float(np.linalg.norm(np.array(xd.pos[:]) - np.array(xh.pos[:]))))
        assert np.linalg.norm(np.array(xd.pos[:]) - np.array(xo.pos[:])) < 1e-3
        n += 1
    assert n >= 30
    print(f"device IMU propagation over {n} scans: prior state vs oracle <= {worst_x:.1e}, prior P (relative) <= {worst_P:.1e}, posterior vs host-propagating builder <= {worst_post:.1e} m")
    assert worst_x < 1e-8 and worst_P < 1e-7 and worst_post < 1e-6
    assert_maps_equal(host.map.dump_map(), dev.map.dump_map(), exact=False, rtol=1e-6, what="device-propagated map")


def test_city_run_with_continuous_eviction(oracle_mod):
    """C3 in small (BASELINE.json configs[2]): drive along a street of scene B at up to 5 m/s with a map capacity far below
    the voxels seen, so that LRU eviction runs in every scan (tens of thousands of victims over the run).
    The map is fed with the oracle's world points: counters and evicted keys per scan and the final map are bit-exact.
    (No free-running comparison here: with so small a map the estimator itself drifts by metres along the street, and two
    runs that differ in the 9th digit end up in different places; the free-running check under eviction is the next test.)"""
    cfg = default_config(max_points_per_scan=8192, map_capacity=8000)
    o = oracle_mod.Oracle(cfg)
    g = HotPath(cfg)
    seq = synth.Sequence(scene=synth.scene_city(), traj=synth.Trajectory(centre=(600.0, 612.0, 1.8), ax=80.0, ay=0.5, period=100.0),
                         sensor=synth.SensorConfig(pts_per_scan=6000))
    evicted_total = 0
    for pk in seq.packages(110):
        c1 = pk.cloud.copy()
        so = o.lio_process(pk.imus, c1, pk.t0, pk.t1)
        _, _, s1 = o.lio_state()
        if s1 < 2 and so.map.n_points == 0:
            continue
        pw, pc = o.dump_world_points(len(c1))
        sg = g.map_build(pw, pc) if so.iters == 0 else g.map_update(pw, pc)
        for f in ("n_ins", "n_touch", "n_created", "n_refit", "refit_points", "n_full", "n_mergeprobe", "n_merge", "n_evicted", "map_size"):
            assert sg[f] == getattr(so.map, f), (pk.index, f)
        assert np.array_equal(o.dump_evicted(), g.dump_evicted()), f"evicted keys differ in scan {pk.index}"
        evicted_total += sg["n_evicted"]
    assert evicted_total > 5000, evicted_total
    assert_maps_equal(o.dump_map(), g.dump_map(), exact=True, what="city map, same world points")


def test_free_running_with_eviction(oracle_mod):
    """scene A with a map capacity below the voxels of the hall: eviction active in most scans, free-running estimators
    (host propagation, device compensation + update) stay within tier 3 and evict the same number of voxels per scan."""
    cfg = default_config(max_points_per_scan=8192, map_capacity=3000)
    o = oracle_mod.Oracle(cfg)
    b = LIOBuilder(cfg)
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=6000))
    worst_p, evicted_total = 0.0, 0
    prev = None
    for pk in seq.packages(80):
        so = o.lio_process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        sb = b.process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        xo, _, s1 = o.lio_state()
        xb, _, s2 = b.state()
        assert s1 == s2
        if s1 == 2 and so.iters:
            worst_p = max(worst_p, float(np.linalg.norm(np.array(xo.pos[:]) - np.array(xb.pos[:]))))
            assert sb.map.n_evicted == so.map.n_evicted and sb.map.map_size == so.map.map_size
            evicted_total += so.map.n_evicted
    assert evicted_total > 1000, evicted_total
    assert worst_p < 1e-3, f"trajectory deviates {worst_p} m"


def test_two_trajectories_on_one_gpu_from_two_threads():
    """Two handles driven concurrently from two host threads: the resident iteration loop of each (CTAs that wait for each other) is
    launched cooperatively, so neither can occupy half of the GPU and starve the other; the results are those of the same scans run alone."""
    import threading
    cfg = default_config(max_points_per_scan=8192)
    seqs = [list(synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=6000), seed=sd).packages(45)) for sd in (11, 12)]

    def run(pkgs, out):
        lio = LIOBuilder(cfg)
        for pk in pkgs:
            lio.process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        x, P, _ = lio.state()
        out.append((bytes(x), P.copy(), lio.map.dump_map()))

    alone = [[], []]
    for k in range(2):
        run(seqs[k], alone[k])
    both = [[], []]
    th = [threading.Thread(target=run, args=(seqs[k], both[k])) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in th), "two concurrent handles did not finish (dead-locked resident kernels?)"
    for k in range(2):
        assert both[k][0][0] == alone[k][0][0] and np.array_equal(both[k][0][1], alone[k][0][1])
        assert_maps_equal(alone[k][0][2], both[k][0][2], exact=True, what="concurrent handle %d" % k)


def test_late_merge_activation_is_detected_and_redone_exactly(monkeypatch):
    """The merge rounds run events in parallel that are far enough apart; an event created on the way (a merge changes planes, the voxels
    around them are re-examined) that should have run BEFORE an event already started is detected through the history of started events,
    and the scan's merge phase is taken back (undo log) and redone in strict event order.  With the default conflict radius (4: one more
    than the exact 3) this has not been seen on any workload; with VMP_MERGE_R=3 it happens about once in 70 scans of the C2 sequence.
    Both settings are exact, so the two runs must agree bit for bit - and the radius-3 run must have gone through the redo."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    wl = dict(name="c2-like", pts=200000, voxel_size=0.25, max_iter=4, capacity=400000)
    pk = bench.make_packages(wl, 0xC0FFEE, 70)
    cfg = default_config(max_points_per_scan=200064, voxel_size=0.25, opti_max_iter=4, map_capacity=400000)
    a = LIOBuilder(cfg)
    monkeypatch.setenv("VMP_MERGE_R", "3")
    b = LIOBuilder(cfg)
    monkeypatch.delenv("VMP_MERGE_R")
    redone = [0, 0]
    for p in pk:
        sa = a.process(p.imus, p.cloud.copy(), p.t0, p.t1)
        sb = b.process(p.imus, p.cloud.copy(), p.t0, p.t1)
        xa, Pa, _ = a.state()
        xb, Pb, _ = b.state()
        assert bytes(xa) == bytes(xb) and np.array_equal(Pa, Pb), f"scan {p.index}"
        assert sa.map.as_dict() == sb.map.as_dict(), f"scan {p.index}"
        if sa.iters:
            redone[0] += (a.map.debug_counters()[2] >> 30) & 1
            redone[1] += (b.map.debug_counters()[2] >> 30) & 1
    assert_maps_equal(a.map.dump_map(), b.map.dump_map(), exact=True, what="conflict radius 4 vs 3")
    assert redone[0] == 0 and redone[1] >= 1, redone
