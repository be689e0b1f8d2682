"""Diagnostic: per-scan counters of the merge simulation (active set, events, rounds, activations) on a bench workload (GPU).
  python tools/merge_stats.py [scans=70] [pts=200000] [voxel=0.25] [max_iter=4] [capacity=400000]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from voxelmapplus_fastlio2_b200.ctypes_defs import default_config  # noqa: E402
from voxelmapplus_fastlio2_b200.lio import LIOBuilder  # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 70
pts = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
voxel = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
max_iter = int(sys.argv[4]) if len(sys.argv) > 4 else 4
cap = int(sys.argv[5]) if len(sys.argv) > 5 else 400000
wl = dict(name="diag", pts=pts, voxel_size=voxel, max_iter=max_iter, capacity=cap)
pk = bench.make_packages(wl, 0xC0FFEE, n_scans)
cfg = default_config(max_points_per_scan=pts + 64, voxel_size=voxel, opti_max_iter=max_iter, map_capacity=cap)
lio = LIOBuilder(cfg)
rows = []
for p in pk:
    st = lio.process(p.imus, p.cloud.copy(), p.t0, p.t1)
    if st.iters == 0:
        continue
    d = lio.map.debug_counters()
    rows.append((p.index, st.iters, st.map.n_touch, st.map.n_full, st.map.n_mergevox, st.map.n_merge, d[0], d[1], d[2] & 0xFFFF, (d[2] >> 16) & 0x3FFF, st.gpu_ms, (d[2] >> 30) & 1))
    if p.index % 3 == 0:
        print("scan %3d iters %d touch %5d full %6d mergevox %5d merges %3d | active0 %4d events %4d react %3d rounds %3d | gpu %.3f ms | serial redo %d" % rows[-1])
a = np.array(rows, float)
s = a[a[:, 0] >= 45]
print("mean (scans >= 45): merges %.1f active0 %.1f events %.1f react %.1f rounds %.1f gpu_ms %.3f; scans redone serially: %d of %d" % (tuple(s[:, k].mean() for k in (5, 6, 7, 8, 9, 10)) + (int(a[:, 11].sum()), len(a))))
