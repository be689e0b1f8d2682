"""SURVEY.md 8(f) row 1, second half: scan_filter.filter() = pcl::VoxelGrid<PointXYZINormal>::filter
(lio_builder.cpp:13-14, 215-219).

CPU: the oracle's restatement against an independent numpy float32 restatement of PCL's published algorithm
(parity unpinned: PCL is not in the image, the reference pins no version).  GPU (`-m gpu`): the device filter through the
C ABI against the oracle, bit for bit (float32 sums in the original point order inside a leaf), stand-alone, teacher-forced
inside a scan, and free-running with the reference's default scan_resolution = 0.1.
"""
import numpy as np
import pytest

from voxelmapplus_fastlio2_b200 import synth
from voxelmapplus_fastlio2_b200.ctypes_defs import default_config

F32 = np.float32


def np_voxel_grid(cloud, leaf):
    """numpy restatement (float32 arithmetic, one operation at a time)."""
    cloud = np.asarray(cloud, F32).reshape(-1, 4)
    inv = F32(1.0) / F32(leaf)
    fin = np.isfinite(cloud[:, :3]).all(axis=1)
    if not fin.any():
        return np.zeros((0, 4), F32)
    mn, mx = cloud[fin, :3].min(axis=0), cloud[fin, :3].max(axis=0)
    d = ((mx - mn) * inv).astype(np.int64) + 1
    if int(d[0]) * int(d[1]) * int(d[2]) > 2 ** 31 - 1:
        return cloud.copy()
    min_b = np.floor(mn * inv).astype(np.int32)
    max_b = np.floor(mx * inv).astype(np.int32)
    div = max_b - min_b + 1
    ijk = (np.floor(cloud[:, :3] * inv) - min_b.astype(F32)).astype(np.int32)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    sel = np.nonzero(fin)[0]
    order = sel[np.argsort(idx[sel], kind="stable")]
    out = []
    start = 0
    sidx = idx[order]
    while start < len(order):
        end = start + 1
        while end < len(order) and sidx[end] == sidx[start]:
            end += 1
        s = np.zeros(4, F32)
        for j in order[start:end]:
            s = s + cloud[j]                       # float32, sequential, original order inside the leaf
        out.append(s / F32(end - start))
        start = end
    return np.array(out, F32).reshape(-1, 4)


def _cloud(seed, n, extent=8.0):
    rng = np.random.Generator(np.random.Philox(key=seed))
    c = np.empty((n, 4), F32)
    c[:, :3] = rng.uniform(-extent, extent, (n, 3)).astype(F32)
    # half of the points on a few planes so that leaves hold several points
    k = n // 2
    c[:k, 2] = (0.02 * rng.normal(size=k)).astype(F32)
    c[:, 3] = np.sort(rng.uniform(0, 100, n)).astype(F32)
    return c


CASES = [("random", lambda: (_cloud(1, 3000), 0.1)), ("coarse", lambda: (_cloud(2, 3000), 0.5)),
         ("one_leaf", lambda: (_cloud(3, 200, extent=0.04) + F32(0.05), 0.5)), ("single", lambda: (_cloud(4, 1), 0.1)),
         ("empty", lambda: (np.zeros((0, 4), F32), 0.1)),
         ("leaf_too_small", lambda: (_cloud(5, 500, extent=30.0), 1e-3))]


def _with_nan(c):
    c = c.copy()
    c[5, 0] = np.nan
    c[17, 2] = np.inf
    return c


@pytest.mark.parametrize("name", [c[0] for c in CASES] + ["non_finite"])
def test_oracle_filter_matches_numpy(oracle_mod, name):
    if name == "non_finite":
        cloud, leaf = _with_nan(_cloud(6, 1000)), 0.2
    else:
        cloud, leaf = dict(CASES)[name]()
    o = oracle_mod.Oracle(default_config(max_points_per_scan=4096))
    got = o.downsample(cloud, leaf)
    want = np_voxel_grid(cloud, leaf)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got, want, equal_nan=True)
    if name == "leaf_too_small":
        assert np.array_equal(got, cloud)               # PCL: "leaf size is too small", output = input
    if name == "one_leaf":
        assert got.shape[0] == 1
    if name in ("random", "coarse"):
        assert 0 < got.shape[0] < cloud.shape[0]
        # every centroid lies in the leaf its points came from, and the point count is conserved by construction
        assert np.isfinite(got).all()


def test_oracle_process_uses_the_filter(oracle_mod):
    """LIOBuilder::process with scan_resolution > 0 (lio_builder.cpp:215-219): the filter input are the leaf centroids of
    the undistorted cloud; MAP_INIT builds from the raw cloud (Q24)."""
    cfg = default_config(max_points_per_scan=4096, scan_resolution=0.25)
    o = oracle_mod.Oracle(cfg)
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=3000))
    seen = 0
    for pk in seq.packages(6):
        cloud = pk.cloud.copy()
        st = o.lio_process(pk.imus, cloud, pk.t0, pk.t1)
        _, _, status = o.lio_state()
        if status == 2 and st.iters > 0:
            lc = o.lidar_cloud()
            assert np.array_equal(lc, np_voxel_grid(cloud, 0.25))       # `cloud` was undistorted in place
            assert st.map.n_points == lc.shape[0] < cloud.shape[0]
            seen += 1
        elif status == 2:
            assert st.map.n_points == cloud.shape[0]                    # map->build on the raw cloud
    assert seen >= 2


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", [c[0] for c in CASES] + ["non_finite", "dense_200k"])
def test_device_filter_bit_exact(oracle_mod, name):
    from voxelmapplus_fastlio2_b200.bindings import HotPath
    if name == "non_finite":
        cloud, leaf = _with_nan(_cloud(6, 1000)), 0.2
    elif name == "dense_200k":
        pk = next(iter(synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=200000)).packages(1, start=30)))
        cloud, leaf = pk.cloud, 0.1
    else:
        cloud, leaf = dict(CASES)[name]()
    cfg = default_config(max_points_per_scan=max(4096, cloud.shape[0] + 64), map_capacity=1000)
    o = oracle_mod.Oracle(cfg)
    g = HotPath(cfg)
    want = o.downsample(cloud, leaf)
    got = g.downsample(cloud, leaf)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got, want, equal_nan=True)
    again = g.downsample(cloud, leaf)                   # scratch is reset between calls
    assert np.array_equal(again, want, equal_nan=True)


@pytest.mark.gpu
def test_scan_on_filtered_cloud_teacher_forced(oracle_mod):
    """The oracle's undistorted cloud goes through the device filter (bit-exact leaf centroids), then through vmp_scan from
    the oracle's prior: iteration counts, valid correspondences and voxel keys of the FILTERED points are bit-exact."""
    from voxelmapplus_fastlio2_b200.bindings import HotPath
    cfg = default_config(max_points_per_scan=8192, scan_resolution=0.1)
    o = oracle_mod.Oracle(cfg)
    g = HotPath(cfg)
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=6000))
    checked = 0
    for pk in seq.packages(12):
        cloud = pk.cloud.copy()
        st = o.lio_process(pk.imus, cloud, pk.t0, pk.t1)
        x_post, _, status = o.lio_state()
        if status == 1:
            continue
        x0, P0 = o.get_prior()
        if st.iters == 0:
            g.first_scan(x0, P0, np.ascontiguousarray(cloud[:, :3]))
            continue
        ds = g.downsample(cloud, cfg.scan_resolution)
        assert np.array_equal(ds, o.lidar_cloud())
        xg, Pg, sg = g.scan(x0, P0, np.ascontiguousarray(ds[:, :3]))
        assert sg.iters == st.iters and list(sg.effect_num[:st.iters]) == list(st.effect_num[:st.iters])
        assert sg.map.n_points == ds.shape[0]
        co, cg = o.dump_correspondences(ds.shape[0]), g.dump_correspondences(ds.shape[0])
        assert np.array_equal(co["keys"], cg["keys"]) and np.array_equal(co["status"], cg["status"])
        assert np.abs(np.array(xg.pos[:]) - np.array(x_post.pos[:])).max() < 1e-8
        checked += 1
    assert checked >= 8


@pytest.mark.gpu
@pytest.mark.parametrize("device_undistort", [True, False])
def test_lio_with_default_scan_resolution(oracle_mod, device_undistort):
    """Free-running LIOBuilder with the reference's default scan_resolution = 0.1: motion compensation (device or host),
    filter on the device, update; tier 3 against the oracle.  A float32-ulp difference of a compensated point can move
    it across a leaf boundary, so leaf counts may differ by a few points; the trajectory may not."""
    from voxelmapplus_fastlio2_b200.lio import LIOBuilder
    cfg = default_config(max_points_per_scan=8192, scan_resolution=0.1)
    o = oracle_mod.Oracle(cfg)
    b = LIOBuilder(cfg, device_undistort=device_undistort)
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=6000))
    worst_p = worst_r = 0.0
    n_checked = 0
    for pk in seq.packages(40):
        c1, c2 = pk.cloud.copy(), pk.cloud.copy()
        so = o.lio_process(pk.imus, c1, pk.t0, pk.t1)
        sb = b.process(pk.imus, c2, pk.t0, pk.t1)
        xo, _, s1 = o.lio_state()
        xb, _, s2 = b.state()
        assert s1 == s2
        if s1 == 2 and so.iters:
            assert abs(int(sb.map.n_points) - int(so.map.n_points)) <= 3
            assert sb.map.n_points < c2.shape[0]                        # the filter really ran
            if device_undistort:
                lc = b.map.lidar_cloud()
                assert lc.shape[0] == sb.map.n_points
            worst_p = max(worst_p, float(np.linalg.norm(np.array(xo.pos[:]) - np.array(xb.pos[:]))))
            worst_r = max(worst_r, synth.rot_angle_deg(np.array(xo.rot[:]).reshape(3, 3), np.array(xb.rot[:]).reshape(3, 3)))
            n_checked += 1
    assert n_checked >= 30
    assert worst_p < 1e-3, f"trajectory deviates {worst_p} m"
    assert worst_r < 1e-2, f"attitude deviates {worst_r} deg"
