"""Run-to-run determinism of the device map update: the same batches through R fresh handles; counters of every batch, the evicted keys
and the final map must be identical bytes (the update is a deterministic function of its input; any difference is a race).
  python tools/determinism_check.py [--points 6000000] [--batch 200000] [--capacity 100000] [--repeats 4]"""
import argparse
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench_map  # noqa: E402
from helpers import assert_maps_equal  # noqa: E402
from voxelmapplus_fastlio2_b200.bindings import HotPath  # noqa: E402
from voxelmapplus_fastlio2_b200.ctypes_defs import default_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=6_000_000)
ap.add_argument("--batch", type=int, default=200_000)
ap.add_argument("--capacity", type=int, default=100_000)
ap.add_argument("--repeats", type=int, default=4)
ap.add_argument("--c3", type=int, default=0, help="instead of the batches: the C3 city drive, this many scans, free-running vmp::LIOBuilder")
a = ap.parse_args()
if a.c3 > 0:
    # every kernel of the path (motion compensation, IEKF, map update with continuous LRU eviction) on the same packages, R times
    import parity_cases as pc
    from voxelmapplus_fastlio2_b200 import synth
    from voxelmapplus_fastlio2_b200.lio import LIOBuilder
    traj = synth.Trajectory(centre=(600.0, 600.0, 1.8), ax=560.0, ay=2.0, period=448.0)
    seq = synth.Sequence(scene=synth.scene_city(pilasters=True), traj=traj, sensor=synth.SensorConfig(pts_per_scan=20000), seed=0xC3, cull=True)
    pk = pc._packages(seq, a.c3 + 3)
    cfg = default_config(max_points_per_scan=20064, map_capacity=a.capacity)
    ref = ref_map = None
    bad = 0
    for r in range(a.repeats):
        lio = LIOBuilder(cfg)
        trace, redo, ev_total = [], 0, 0
        for p in pk:
            st = lio.process(p.imus, p.cloud.copy(), p.t0, p.t1)
            ev = lio.map.dump_evicted() if st.map.n_evicted else np.zeros((0, 3), np.int64)
            ev_total += int(st.map.n_evicted)
            redo += (lio.map.debug_counters()[2] >> 30) & 1
            x, P, _ = lio.state()
            trace.append((st.iters, tuple(st.effect_num[:]), tuple(sorted(st.map.as_dict().items())), hashlib.sha1(ev.tobytes()).hexdigest(),
                          hashlib.sha1(bytes(x) + P.tobytes()).hexdigest()))
        mp_ = lio.map.dump_map()
        lio.close()
        if ref is None:
            ref, ref_map = trace, mp_
            print(f"run 0: {len(pk)} packages, {len(mp_)} voxels, evicted {ev_total}, scans redone serially {redo}")
            continue
        diff = [k for k in range(len(trace)) if trace[k] != ref[k]]
        msg = "identical" if not diff else "DIFFERS first at package %d: %s vs %s" % (diff[0], trace[diff[0]][:3], ref[diff[0]][:3])
        try:
            assert_maps_equal(ref_map, mp_, exact=True, what="final map")
        except AssertionError as e:
            msg += "; " + str(e)[:300]
            diff.append(-1)
        print(f"run {r}: {msg} (scans redone serially {redo})")
        bad += bool(diff)
    print("DETERMINISTIC" if not bad else f"NON-DETERMINISTIC in {bad} of {a.repeats - 1} repeats")
    sys.exit(1 if bad else 0)
nb = max(3, a.points // a.batch)
batches = bench_map.make_batches(nb, a.batch)
cfg = default_config(max_points_per_scan=a.batch + 64, map_capacity=a.capacity)
ref = ref_map = None
bad = 0
for r in range(a.repeats):
    g = HotPath(cfg)
    trace = []
    redo = 0
    for k, (p, c) in enumerate(batches):
        st = g.map_update(p, c) if k else g.map_build(p, c)
        ev = g.dump_evicted() if st["n_evicted"] else np.zeros((0, 3), np.int64)
        redo += (g.debug_counters()[2] >> 30) & 1
        trace.append((tuple(sorted(st.items())), hashlib.sha1(ev.tobytes()).hexdigest()))
    mp = g.dump_map()
    g.close()
    if ref is None:
        ref, ref_map = trace, mp
        print(f"run 0: {nb} batches, {len(mp)} voxels, merges {sum(dict(t[0])['n_merge'] for t in trace[:-1])}, batches redone serially {redo}")
    else:
        diff = [k for k in range(len(trace)) if trace[k] != ref[k]]
        msg = "identical" if not diff else "DIFFERS first at batch %d" % diff[0]
        try:
            assert_maps_equal(ref_map, mp, exact=True, what="final map")       # (group ids are compared as a partition: only equality matters, Q22)
        except AssertionError as e:
            msg += "; " + str(e)[:300]
            diff.append(-1)
        print(f"run {r}: {msg} (batches redone serially {redo})")
        bad += bool(diff)
print("DETERMINISTIC" if not bad else f"NON-DETERMINISTIC in {bad} of {a.repeats - 1} repeats")
sys.exit(1 if bad else 0)
