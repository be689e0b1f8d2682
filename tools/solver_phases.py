"""Diagnostic: cycles of the solver CTA's phases (k_measure block 0) on the C2 bench workload, iteration VMP_DEBUG_ITER (default 1).
  python tools/solver_phases.py [scans=60]"""
import os
import sys

import numpy as np

os.environ.setdefault("VMP_DEBUG_ITER", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from voxelmapplus_fastlio2_b200.ctypes_defs import default_config  # noqa: E402
from voxelmapplus_fastlio2_b200.lio import LIOBuilder  # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 60
wl = dict(name="diag", pts=200000, voxel_size=0.25, max_iter=4, capacity=400000)
pk = bench.make_packages(wl, 0xC0FFEE, n_scans)
cfg = default_config(max_points_per_scan=200064, voxel_size=0.25, opti_max_iter=4, map_capacity=400000)
lio = LIOBuilder(cfg)
rows = []
for p in pk:
    st = lio.process(p.imus, p.cloud.copy(), p.t0, p.t1)
    if st.iters < 2:
        continue
    d = lio.map.debug_counters()
    rows.append([p.index, st.iters] + [int(x) for x in d[3:8]])
a = np.array(rows, float)
s = a[a[:, 0] >= 40]
print("iteration %s, mean cycles over %d scans: A(setup) %.0f  W(wait for the measurement CTAs) %.0f  reduce+solve %.0f  boxplus %.0f  posterior(last) %.0f"
      % ((os.environ["VMP_DEBUG_ITER"], len(s)) + tuple(s[:, k].mean() for k in (2, 3, 4, 5, 6))))
