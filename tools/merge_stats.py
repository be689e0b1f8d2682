"""Diagnostic: per-scan counters of the merge simulation and per-kernel times on the C1 workload (GPU)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelmapplus_fastlio2_b200 import synth  # noqa: E402
from voxelmapplus_fastlio2_b200.ctypes_defs import default_config  # noqa: E402
from voxelmapplus_fastlio2_b200.lio import LIOBuilder  # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 60
pts = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
cfg = default_config(max_points_per_scan=pts + 64)
lio = LIOBuilder(cfg)
seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=pts))
rows = []
for pk in seq.packages(n_scans):
    st = lio.process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
    if st.iters == 0:
        continue
    d = lio.map.debug_counters()
    rows.append((pk.index, st.iters, st.map.n_touch, st.map.n_full, st.map.n_merge, d[0], d[1], d[2], st.gpu_ms, st.host_ms))
    if pk.index % 5 == 0:
        print("scan %3d iters %d touch %5d full %6d merges %3d | active0 %4d events %4d react %3d | gpu %.3f ms host %.3f ms" % rows[-1])
print("solver CTA cycles (last scan): A boxminus|J|A^-1 %d, wait for the measurement %d, reduce + DxD algebra %d, boxplus %d, posterior %d" % tuple(d[3:8]))
a = np.array(rows, float)
print("mean: merges %.1f active0 %.1f events %.1f react %.1f gpu_ms %.3f" % (a[:, 4].mean(), a[:, 5].mean(), a[:, 6].mean(), a[:, 7].mean(), a[10:, 8].mean()))

# per-kernel CUDA-event times (scans run kernel by kernel instead of the graph)
lio.map.profile_enable(True)
lio.map.profile_reset()
seq2 = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=pts))
n_prof = 0
for pk in seq.packages(n_scans + 20):
    if pk.index < n_scans:
        continue
    lio.process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
    n_prof += 1
prof = lio.map.profile_read()
tot = sum(v[0] for v in prof.values())
for name, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    if cnt:
        print("%-18s %6.1f us/scan  %5.2f us/launch  %4.1f launches/scan  %4.1f%%" % (name, 1e3 * ms / n_prof, 1e3 * ms / cnt, cnt / n_prof, 100 * ms / tot))
print("total %.1f us/scan over %d scans" % (1e3 * tot / n_prof, n_prof))
