"""Longer free-running comparison of the independent Python implementation (tests/lio_pyref.py) with the C++ oracle: python tools/pin_long.py <pts_per_scan> <scans>."""
import sys, time; import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle_py
from lio_pyref import LioPy
from voxelmapplus_fastlio2_b200 import synth
from voxelmapplus_fastlio2_b200.ctypes_defs import default_config
oracle_py.build()
cfg = default_config(max_points_per_scan=4096, map_capacity=100000)
o = oracle_py.Oracle(cfg); py = LioPy(cfg)
seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=int(sys.argv[1])))
t0=time.time(); worst=dict(pos=0.0,rot=0.0,P=0.0); upd=0; eq=0; travelled=0.0; last=None
for pk in seq.packages(int(sys.argv[2])):
    st = o.lio_process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
    py.process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
    xo,Po,status = o.lio_state()
    if status<2 or st.iters==0: continue
    d=xo.as_dict()
    assert st.iters==py.iters
    eq += list(st.effect_num[:st.iters])==py.effect
    worst["pos"]=max(worst["pos"],np.abs(d["pos"]-py.x["pos"]).max()); worst["rot"]=max(worst["rot"],np.abs(d["rot"].reshape(3,3)-py.x["rot"]).max()); worst["P"]=max(worst["P"],np.abs(Po-py.P).max()/np.abs(Po).max())
    if last is not None: travelled += np.linalg.norm(d["pos"]-last)
    last=d["pos"].copy(); upd+=1
print(f"python LIO vs oracle, free-running, {sys.argv[1]} pts/scan: {upd} updates, {travelled:.1f} m travelled, iterations equal in all, effect_num equal in {eq}, worst {worst}, map {len(py.map.feat)} voxels == {o.map_size()} ({time.time()-t0:.0f} s)")
