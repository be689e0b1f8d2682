"""Generates the committed golden vectors under tests/golden/ from the CPU oracle.

The reference ships no golden vectors and cannot be compiled in this image (Eigen/Sophus/PCL/ROS absent),
so these fixtures pin the ORACLE (against accidental change) and give the GPU tests a fixed target that does
not depend on the oracle being re-run.  Regenerate with:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle_py  # noqa: E402
from voxelmapplus_fastlio2_b200 import synth  # noqa: E402
from voxelmapplus_fastlio2_b200.ctypes_defs import default_config  # noqa: E402

MAP_CFG = dict(max_point_thresh=30, update_size_thresh=5, map_capacity=500, max_points_per_scan=2048)
LIO_CFG = dict(max_points_per_scan=2048)
LIO_PTS, LIO_SCANS = 1500, 12


def map_workload():
    """8 scans of 900 float32-rounded points: two perpendicular walls + floor + clutter, drifting 0.7 m per scan so
    that voxels fill, close, merge and (capacity 500) get evicted.  Deterministic, no RNG library involved."""
    scans = []
    for s in range(8):
        i = np.arange(900, dtype=np.float64)
        a = (i * 0.6180339887498949 + 0.13 * s) % 1.0
        b = (i * 0.7548776662466927 + 0.29 * s) % 1.0
        c = (i * 0.5698402909980532 + 0.41 * s) % 1.0
        kind = (np.arange(900) % 4)
        x0 = 0.7 * s
        p = np.empty((900, 3))
        w = 0.004 * np.sin(37.0 * i + s)                      # few-mm roughness
        m = kind == 0; p[m] = np.stack([x0 + 4 * a[m], 0.26 + w[m], 2 * b[m]], 1)
        m = kind == 1; p[m] = np.stack([x0 + 0.13 + w[m], 3 * a[m], 2 * b[m]], 1)
        m = kind == 2; p[m] = np.stack([x0 + 4 * a[m], 3 * b[m], 0.07 + w[m]], 1)
        m = kind == 3; p[m] = np.stack([x0 + 5 * a[m] - 0.5, 4 * b[m] - 0.5, 3 * c[m]], 1)
        p = p.astype(np.float32).astype(np.float64)
        cov = np.zeros((900, 9))
        cov[:, 0] = 1e-4 * (1 + (np.arange(900) % 7) / 10.0)
        cov[:, 4] = 1.5e-4
        cov[:, 8] = 0.8e-4 * (1 + (np.arange(900) % 3) / 5.0)
        cov[:, 1] = 1e-6; cov[:, 3] = 1.0000001e-6                 # pv.cov is not exactly symmetric in the reference either
        scans.append((p, cov))
    return scans


DS_LEAF = 0.2


def downsample_cloud():
    """3000 points (x y z curvature, float32) on two walls, a floor and clutter, incl. two non-finite points; no RNG library."""
    i = np.arange(3000, dtype=np.float64)
    a = (i * 0.6180339887498949) % 1.0
    b = (i * 0.7548776662466927) % 1.0
    c = (i * 0.5698402909980532) % 1.0
    kind = np.arange(3000) % 4
    p = np.empty((3000, 4))
    w = 0.004 * np.sin(37.0 * i)
    m = kind == 0; p[m, :3] = np.stack([6 * a[m] - 3, 2.26 + w[m], 2 * b[m]], 1)
    m = kind == 1; p[m, :3] = np.stack([-3.13 + w[m], 5 * a[m] - 2.5, 2 * b[m]], 1)
    m = kind == 2; p[m, :3] = np.stack([6 * a[m] - 3, 5 * b[m] - 2.5, -0.93 + w[m]], 1)
    m = kind == 3; p[m, :3] = np.stack([7 * a[m] - 3.5, 6 * b[m] - 3, 3 * c[m] - 1], 1)
    p[:, 3] = i / 30.0
    p = p.astype(np.float32)
    p[11, 0] = np.nan
    p[1234, 2] = np.inf
    return p


def lio_packages():
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=LIO_PTS), seed=20261017)
    return list(seq.packages(LIO_SCANS))


def input_digest(pkgs):
    h = hashlib.sha256()
    for pk in pkgs:
        h.update(pk.cloud.tobytes()); h.update(pk.imus.tobytes())
    return h.hexdigest()


def main():
    oracle_py.build()
    # ---- map golden
    o = oracle_py.Oracle(default_config(**MAP_CFG))
    counters, evicted = [], []
    for s, (p, c) in enumerate(map_workload()):
        st = o.map_update(p, c) if s else o.map_build(p, c)
        counters.append([st[k] for k in sorted(st)])
        evicted.append(o.dump_evicted())
    d = o.dump_map()
    np.savez_compressed(os.path.join(HERE, "map_golden.npz"), counters=np.array(counters, np.int64),
                        counter_names=np.array(sorted(st)), evicted=np.concatenate(evicted), evicted_count=np.array([len(e) for e in evicted]),
                        key=d["key"], flags=d["flags"], n=d["n"], n_temp=d["n_temp"], newly=d["newly_add_point"],
                        mean=d["mean"], norm=d["norm"], cov=d["cov"], ppt=d["ppt"])
    # ---- LIO golden
    pk = lio_packages()
    o = oracle_py.Oracle(default_config(**LIO_CFG))
    pos, rot, iters, eff, H_last, b_last = [], [], [], [], None, None
    for p in pk:
        st = o.lio_process(p.imus, p.cloud.copy(), p.t0, p.t1)
        x, P, status = o.lio_state()
        pos.append(list(x.pos)); rot.append(list(x.rot)); iters.append(st.iters); eff.append(list(st.effect_num))
        if st.iters:
            H_last, b_last = o.get_iter_Hb(0)
    np.savez_compressed(os.path.join(HERE, "lio_golden.npz"), pos=np.array(pos), rot=np.array(rot), iters=np.array(iters),
                        effect=np.array(eff), H_last=H_last, b_last=b_last, P_last=P, digest=np.array(input_digest(pk)),
                        map_size=np.array(o.map_size()))
    # ---- scan filter golden (pcl::VoxelGrid restatement, lio_builder.cpp:215-219)
    ds = oracle_py.Oracle(default_config(max_points_per_scan=4096)).downsample(downsample_cloud(), DS_LEAF)
    np.savez_compressed(os.path.join(HERE, "downsample_golden.npz"), out=ds, leaf=np.array(DS_LEAF))
    print("written", os.listdir(HERE))


if __name__ == "__main__":
    main()
