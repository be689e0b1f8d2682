"""Independent pins of the oracle (parity is otherwise unpinned: the reference has no vectors and cannot be built here).

updatePlane end to end (voxel_map.cpp:97-136) - covariance of the stored points, eigen-decomposition, plane test, normal with its sign
rule, and the 6x6 plane covariance sum_i J_i Sigma_i J_i^T - recomputed with numpy / LAPACK (`numpy.linalg.eigh`) straight from the
formulas of the reference, for every plane voxel of an oracle map.  LAPACK's eigenvectors come with arbitrary signs: the two diagonal
3x3 blocks of the plane covariance are invariant under them, the off-diagonal blocks flip with the sign of the normal's eigenvector, so
those are compared up to that one common sign.  (A misreading of Eigen's conventions - eigenvalue order, which eigenvector is the
normal, the role of lambda_0 - lambda_m - shows up here; tests/test_oracle_math.py only pins the eigen-solver itself.)"""
import numpy as np

from helpers import F_PLANE, plane_cloud, random_cov
from voxelmapplus_fastlio2_b200.ctypes_defs import default_config


def _plane_fit_numpy(P, S):
    """P: n x 3 points, S: n x 3 x 3 covariances -> mean, normal (sign rule applied), 6x6 covariance, smallest eigenvalue, raw normal"""
    n = len(P)
    mean = np.zeros(3)
    for i, p in enumerate(P):                       # the running mean of addToPlane (voxel_map.cpp:29-34)
        mean = mean + (p - mean) / (i + 1.0)
    ppt = sum(np.outer(p, p) for p in P)
    cov = ppt / n - np.outer(mean, mean)
    w, V = np.linalg.eigh(cov)                      # ascending, like SelfAdjointEigenSolver
    nrm = V[:, 0]
    C = np.zeros((6, 6))
    for p, s in zip(P, S):
        F = np.zeros((3, 3))
        for m in (1, 2):
            F[m] = (p - mean) / (n * (w[0] - w[m])) @ (np.outer(V[:, m], nrm) + np.outer(nrm, V[:, m]))
        J = np.vstack([V @ F, np.eye(3) / n])
        C += J @ s @ J.T
    raw = nrm.copy()
    if -(mean @ nrm) < 0:
        nrm = -nrm
    return mean, nrm, C, w[0], raw


def test_update_plane_matches_numpy_eigh(oracle_mod):
    rng = np.random.Generator(np.random.Philox(key=31))
    cfg = default_config(max_points_per_scan=8192, voxel_size=0.5)
    o = oracle_mod.Oracle(cfg)
    # a few tilted planes crossing many voxels, 12 .. 60 points per voxel
    pts = np.concatenate([plane_cloud(rng, 2500, (0.3, 0.2, 1.0), (1, 0.2, 0.1), (0, 1, 0.3), (4.0, 3.0), 0.004),
                          plane_cloud(rng, 2500, (-3.0, 1.0, 0.2), (0, 1, 0), (0.1, 0, 1), (3.0, 3.0), 0.004),
                          plane_cloud(rng, 1500, (1.0, -4.0, -0.5), (1, 0, 0), (0, 1, 0.05), (3.0, 3.0), 0.004)])
    pts = pts.astype(np.float32).astype(np.float64)
    covs = random_cov(rng, len(pts), scale=1e-4)
    o.map_build(pts, covs)
    d = o.dump_map()
    keys = np.floor(pts / cfg.voxel_size).astype(np.int64)
    checked = 0
    worst = dict(mean=0.0, norm=0.0, diag=0.0, off=0.0)
    for v in d:
        if not (v["flags"] & F_PLANE):
            continue
        idx = np.nonzero((keys == v["key"]).all(axis=1))[0]           # build(): points in input order
        assert len(idx) == v["n"] >= cfg.update_size_thresh
        mean, nrm, C, lam0, raw = _plane_fit_numpy(pts[idx], covs[idx].reshape(-1, 3, 3))
        assert lam0 <= cfg.plane_thresh
        Co = v["cov"].reshape(6, 6)
        sc = max(np.abs(C).max(), 1e-300)
        worst["mean"] = max(worst["mean"], np.abs(v["mean"] - mean).max())
        worst["norm"] = max(worst["norm"], np.abs(v["norm"] - nrm).max())
        worst["diag"] = max(worst["diag"], np.abs(Co[:3, :3] - C[:3, :3]).max() / sc, np.abs(Co[3:, 3:] - C[3:, 3:]).max() / sc)
        # off-diagonal blocks: equal up to the sign of the raw eigenvector (the oracle's solver and LAPACK choose independently)
        off = min(np.abs(Co[:3, 3:] - C[:3, 3:]).max(), np.abs(Co[:3, 3:] + C[:3, 3:]).max()) / sc
        worst["off"] = max(worst["off"], off)
        assert np.abs(Co[3:, :3] - Co[:3, 3:].T).max() <= 1e-9 * sc
        checked += 1
    assert checked >= 40, checked
    print(f"updatePlane vs numpy.linalg.eigh over {checked} plane voxels: {worst}")
    assert worst["mean"] < 1e-12 and worst["norm"] < 1e-8 and worst["diag"] < 1e-8 and worst["off"] < 1e-8, worst


# --------------------------------------------------------------------------- the measurement model, end to end
def _hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def _rot(axis_angle):
    from scipy.spatial.transform import Rotation
    return Rotation.from_rotvec(axis_angle).as_matrix()


def _body_cov_numpy(p, range_cov=0.04, angle_cov=0.1):
    """commons.cpp:18-45 (tests/test_oracle_math.py::test_calc_body_cov pins the oracle's version against the same formulas)"""
    q = p.copy()
    if q[2] == 0:
        q[2] = 0.001
    r = np.linalg.norm(q); d = q / r
    b1 = np.array([1, 1, -(d[0] + d[1]) / d[2]]); b1 /= np.linalg.norm(b1)
    b2 = np.cross(b1, d); b2 /= np.linalg.norm(b2)
    A = r * _hat(d) @ np.stack([b1, b2], 1)
    return q, np.outer(d, d) * range_cov ** 2 + A @ (np.eye(2) * np.sin(angle_cov * 0.017453293) ** 2) @ A.T


def _measure_numpy(planes, voxel_size, x, P, pts, estimate_ext):
    """LIOBuilder::sharedUpdateFunc + VoxelMap::buildResidual (lio_builder.cpp:250-311, voxel_map.cpp:258-276) written from the reference's
    formulas with numpy matrices; plane_cov is the zero matrix the reference never assigns (Q1).  planes: {key tuple: (is_plane, mean, norm)}."""
    R, Rext, pos, pext = x["rot"], x["rot_ext"], x["pos"], x["pos_ext"]
    r_wl = R @ Rext
    p_wl = R @ pext + pos
    H = np.zeros((12, 12)); b = np.zeros(12)
    status = np.zeros(len(pts), np.uint8); res = np.zeros(len(pts)); keys = np.zeros((len(pts), 3), np.int64)
    eff = 0
    for i, p in enumerate(pts):
        pl, cl = _body_cov_numpy(p.astype(np.float64))
        pw = r_wl @ pl + p_wl
        cw = r_wl @ cl @ r_wl.T + _hat(pl) @ P[3:6, 3:6] @ _hat(pl).T + P[0:3, 0:3]
        key = tuple(int(k) for k in np.floor(pw / voxel_size))
        keys[i] = key
        if key not in planes:
            continue                                            # fresh records: is_valid = false
        status[i] |= 1
        is_plane, mean, nrm = planes[key]
        if not is_plane:
            continue
        status[i] |= 2
        r = nrm @ (pw - mean)
        res[i] = r
        if not abs(r) < 3.0 * np.sqrt(nrm @ cw @ nrm):
            continue
        status[i] |= 4
        eff += 1
        r_cov = nrm @ r_wl @ cl @ r_wl.T @ nrm
        w = 5000.0 if r_cov < 0.0002 else 1.0 / r_cov
        J = np.zeros(12)
        J[0:3] = nrm
        J[3:6] = -nrm @ R @ _hat(Rext @ pl + pext)
        if estimate_ext:
            J[6:9] = -nrm @ r_wl @ _hat(pl)
            J[9:12] = nrm @ R
        H += np.outer(J, J) * w
        b += J * w * r
    return H, b, eff, keys, status, res


def _measurement_case(oracle_mod, estimate_ext, seed):
    from voxelmapplus_fastlio2_b200.ctypes_defs import VmpState
    rng = np.random.Generator(np.random.Philox(key=seed))
    cfg = default_config(max_points_per_scan=8192, voxel_size=0.5, estimate_ext=1 if estimate_ext else 0)
    o = oracle_mod.Oracle(cfg)
    # a map of tilted planes + clutter (voxels that exist but are no planes)
    world = np.concatenate([plane_cloud(rng, 3000, (0.3, 0.2, 1.0), (1, 0.2, 0.1), (0, 1, 0.3), (4.0, 3.0), 0.004),
                            plane_cloud(rng, 3000, (-3.0, 1.0, 0.2), (0, 1, 0), (0.1, 0, 1), (3.0, 3.0), 0.004),
                            rng.uniform(-1.0, 0.0, (300, 3)) + np.array([0.0, -2.0, 0.0])]).astype(np.float32).astype(np.float64)
    o.map_build(world, random_cov(rng, len(world), scale=1e-4))
    planes = {tuple(int(k) for k in v["key"]): (bool(v["flags"] & F_PLANE), v["mean"].copy(), v["norm"].copy()) for v in o.dump_map()}
    # a state away from the identity in every block the model reads, and a dense P
    x = VmpState.identity()
    R, Rext = _rot(rng.normal(0, 0.05, 3)), _rot(rng.normal(0, 0.03, 3))
    x.rot[:] = R.ravel(); x.rot_ext[:] = Rext.ravel()
    x.pos[:] = rng.normal(0, 0.05, 3); x.pos_ext[:] = rng.normal(0, 0.05, 3)
    A = rng.normal(0, 1, (23, 23))
    P = (A @ A.T) * 1e-5 + np.eye(23) * 1e-5
    # the scan: new samples of the same surfaces (+ clutter, + points off the map), expressed in the lidar frame
    scan_w = np.concatenate([plane_cloud(rng, 1500, (0.3, 0.2, 1.0), (1, 0.2, 0.1), (0, 1, 0.3), (4.0, 3.0), 0.02),
                             plane_cloud(rng, 1500, (-3.0, 1.0, 0.2), (0, 1, 0), (0.1, 0, 1), (3.0, 3.0), 0.02),
                             plane_cloud(rng, 400, (0.3, 0.2, 1.0), (1, 0.2, 0.1), (0, 1, 0.3), (4.0, 3.0), 0.12),      # outliers for the gate
                             rng.uniform(-1.0, 0.0, (200, 3)) + np.array([0.0, -2.0, 0.0]), rng.uniform(20, 30, (100, 3))])
    r_wl, p_wl = R @ Rext, R @ np.array(x.pos_ext[:]) + np.array(x.pos[:])
    pts = ((scan_w - p_wl) @ r_wl).astype(np.float32)            # r_wl^T (p - p_wl)
    o.set_scan(pts)
    H, b, eff = o.measure(x, P)
    c = o.dump_correspondences(len(pts))
    xd = dict(rot=R, rot_ext=Rext, pos=np.array(x.pos[:]), pos_ext=np.array(x.pos_ext[:]))
    return (H, b, eff, c), _measure_numpy(planes, cfg.voxel_size, xd, P, pts, estimate_ext)


def test_measurement_model_matches_numpy_restatement(oracle_mod):
    """The oracle's sharedUpdateFunc / buildResidual against a numpy restatement written from the reference's formulas (matrix products by
    numpy, rotations by scipy): voxel keys, found / plane / valid flags and effect_num identical, residuals, H and b to 1e-9 relative.
    A misreading of the model - which rotation multiplies which hat matrix, the sign of the Jacobian blocks, the gate's covariance, the
    weight clamp - would show here, independently of the CUDA-vs-oracle tests."""
    for estimate_ext in (False, True):
        for seed in (41, 42):
            (H, b, eff, c), (Hn, bn, effn, keys, status, res) = _measurement_case(oracle_mod, estimate_ext, seed)
            assert eff == effn and eff > 1000
            assert np.array_equal(c["keys"], keys)
            assert np.array_equal(c["status"], status)
            assert (status == 1).any() and (status == 0).any() and (status == 3).any()      # non-plane voxel, off the map, gated out
            v = (status & 4) != 0
            np.testing.assert_allclose(c["residual"][v], res[v], rtol=1e-9, atol=1e-12)
            H = np.asarray(H).reshape(12, 12); b = np.asarray(b).ravel()
            np.testing.assert_allclose(H, Hn, rtol=1e-9, atol=1e-9 * np.abs(Hn).max())
            np.testing.assert_allclose(b, bn, rtol=1e-9, atol=1e-9 * np.abs(bn).max())
            if not estimate_ext:
                assert not H[6:, :].any() and not H[:, 6:].any()


# --------------------------------------------------------------------------- the whole voxel map, second reading
def _compare_with_pyref(o, py, what):
    d = o.dump_map()
    keys = [tuple(k) for k in d["key"].tolist()]
    assert len(keys) == len(py.feat), what
    order = list(py.cache.keys())
    assert keys == order[::-1] or keys == order, f"{what}: LRU order differs"
    F_INIT, F_UPD, F_MERGED = 1, 4, 8
    worst = dict(mean=0.0, norm=0.0, cov=0.0)
    groups_o, groups_p = {}, {}
    for v, k in zip(d, keys):
        g = py.feat[k]
        fl = (F_INIT if g.is_init else 0) | (F_PLANE if g.is_plane else 0) | (F_UPD if g.update_enable else 0) | (F_MERGED if g.merged else 0)
        assert int(v["flags"]) == fl, f"{what}: flags of {k}: {int(v['flags'])} vs {fl}"
        assert int(v["n"]) == g.n and int(v["n_temp"]) == len(g.temp) and int(v["newly_add_point"]) == g.newly, f"{what}: counts of {k}"
        groups_o.setdefault(int(v["group"]), []).append(k); groups_p.setdefault(g.group_id, []).append(k)
        worst["mean"] = max(worst["mean"], np.abs(v["mean"] - g.mean).max())
        np.testing.assert_allclose(v["ppt"].reshape(3, 3), g.ppt, rtol=1e-13, atol=0)
        if g.is_plane:
            worst["norm"] = max(worst["norm"], np.abs(v["norm"] - g.norm).max())
            Co, C = v["cov"].reshape(6, 6), g.cov
            sc = max(np.abs(C).max(), 1e-300)
            # the diagonal blocks do not depend on the eigenvectors' signs (the off-diagonal ones flip with the raw normal's, refit by refit)
            worst["cov"] = max(worst["cov"], np.abs(Co[:3, :3] - C[:3, :3]).max() / sc, np.abs(Co[3:, 3:] - C[3:, 3:]).max() / sc)
    assert sorted(sorted(v) for v in groups_o.values()) == sorted(sorted(v) for v in groups_p.values()), f"{what}: merge groups differ"
    assert worst["mean"] < 1e-12 and worst["norm"] < 1e-7 and worst["cov"] < 1e-6, (what, worst)
    return worst


def test_voxel_map_matches_python_second_reading(oracle_mod):
    """VoxelMap::build / update / pushPoint / addToPlane / updatePlane / merge / the LRU cache: the C++ oracle against tests/voxelmap_pyref.py,
    a plain-Python reading of voxel_map.cpp with LAPACK's eigen-solver.  Identical: keys, LRU order, evicted keys per scan, init / plane /
    update_enable / merged flags, point counts, refit counters, merge groups; means to 1e-12, normals to 1e-7, the sign-independent blocks
    of the plane covariance to 1e-6 relative - through fills, refits, closes, merges and evictions."""
    from helpers import wall_workload
    from voxelmap_pyref import VoxelMapPy
    merges = evictions = 0
    for seed, cap, mk_work in ((1, 100000, None), (2, 100000, None), (5, 400, "moving")):
        cfg = default_config(max_point_thresh=30 if cap > 400 else 20, update_size_thresh=5, map_capacity=cap, max_points_per_scan=4096)
        o = oracle_mod.Oracle(cfg)
        py = VoxelMapPy(cfg.max_point_thresh, cfg.update_size_thresh, cfg.plane_thresh, cfg.voxel_size, cfg.map_capacity)
        if mk_work is None:
            work = wall_workload(seed)
        else:
            rng = np.random.Generator(np.random.Philox(key=seed))
            work = []
            for s in range(25):                     # a corridor that advances 1 m per scan: old voxels fall out of the map
                x0 = 1.0 * s
                p = np.concatenate([np.stack([rng.uniform(x0, x0 + 6, 500), rng.normal(0.25, 0.01, 500), rng.uniform(0, 2, 500)], 1),
                                    np.stack([rng.uniform(x0, x0 + 6, 500), rng.uniform(0, 3, 500), rng.normal(0.1, 0.01, 500)], 1)])
                p = p[rng.permutation(len(p))].astype(np.float32).astype(np.float64)
                work.append((p, np.tile((np.eye(3) * 1e-4).reshape(1, 9), (len(p), 1))))
        for s, (p, c) in enumerate(work):
            c3 = c.reshape(-1, 3, 3)
            if s == 0:
                so = o.map_build(p, c); py.build(p, c3)
            else:
                so = o.map_update(p, c); py.update(p, c3)
            merges += so["n_merge"]
            ev = [tuple(k) for k in o.dump_evicted().tolist()]
            assert ev == py.evicted, f"seed {seed} scan {s}: evicted keys differ"
            evictions += len(ev)
            if s % 4 == 3 or s == len(work) - 1:
                _compare_with_pyref(o, py, f"seed {seed} scan {s}")
    assert merges > 0 and evictions > 100, (merges, evictions)


# --------------------------------------------------------------------------- IMU initialisation + motion compensation
def _so3_exp(v):
    from scipy.spatial.transform import Rotation
    return Rotation.from_rotvec(v).as_matrix()


class _LioNumpy:
    """initializeImu + undistortCloud (lio_builder.cpp:28-153) restated with numpy / scipy from the reference's source: IMU cache, mean
    acceleration -> gravity norm, gravity alignment, the mid-point propagation of pos / rot / vel per IMU step, the pose list, and the backward
    loop that compensates every point into the frame of the scan's end.  The filter's covariance is not needed for any of it."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.cache = []
        self.last_acc = np.zeros(3); self.last_gyro = np.zeros(3)
        self.last_end = 0.0

    def initialize(self, imus, t_end):
        self.cache += [(np.array(i["acc"]), np.array(i["gyro"]), float(i["timestamp"])) for i in imus]
        if len(self.cache) < self.cfg.imu_init_num:
            return None
        acc_mean = sum(c[0] for c in self.cache) / float(len(self.cache))
        gyro_mean = sum(c[1] for c in self.cache) / float(len(self.cache))
        self.gravity_norm = np.linalg.norm(acc_mean)
        x = dict(pos=np.zeros(3), vel=np.zeros(3), ba=np.zeros(3), bg=gyro_mean, rot=np.eye(3),
                 rot_ext=np.array(self.cfg.r_il[:]).reshape(3, 3), pos_ext=np.array(self.cfg.p_il[:]))
        assert self.cfg.gravity_align
        a, b = -acc_mean / np.linalg.norm(acc_mean), np.array([0.0, 0.0, -1.0])          # Quaterniond::FromTwoVectors(a, b): the shortest rotation a -> b
        axis = np.cross(a, b)
        s, c = np.linalg.norm(axis), a @ b
        x["rot"] = _so3_exp(axis / s * np.arctan2(s, c)) if s > 1e-12 else np.eye(3)
        x["g"] = np.array([0.0, 0.0, -1.0]) * 9.81                                       # State::initG: direction scaled to 9.81 (ieskf.h)
        self.last_imu = self.cache[-1]
        self.last_end = t_end
        return x

    def undistort(self, x, imus, cloud, t0, t1):
        """x: state at the start (dict of arrays; pos / rot / vel are advanced in place), cloud: n x 4 float32 (x y z, time offset in ms), sorted."""
        cache = [self.last_imu] + [(np.array(i["acc"]), np.array(i["gyro"]), float(i["timestamp"])) for i in imus]
        poses = [(0.0, self.last_acc.copy(), self.last_gyro.copy(), x["vel"].copy(), x["pos"].copy(), x["rot"].copy())]
        acc = gyro = None

        def predict(acc, gyro, dt):                                  # IESKF::predict, state part (ieskf.cpp:101-108)
            w, a = gyro - x["bg"], acc - x["ba"]
            pos = x["pos"] + x["vel"] * dt
            vel = x["vel"] + (x["rot"] @ a + x["g"]) * dt
            x["rot"] = x["rot"] @ _so3_exp(w * dt)
            x["pos"], x["vel"] = pos, vel

        for head, tail in zip(cache[:-1], cache[1:]):
            if tail[2] < self.last_end:
                continue
            gyro = 0.5 * (head[1] + tail[1])
            acc = 0.5 * (head[0] + tail[0]) * 9.81 / self.gravity_norm
            dt = tail[2] - self.last_end if head[2] < self.last_end else tail[2] - head[2]
            predict(acc, gyro, dt)
            self.last_gyro = gyro - x["bg"]
            self.last_acc = x["rot"] @ (acc - x["ba"]) + x["g"]
            poses.append((tail[2] - t0, self.last_acc.copy(), self.last_gyro.copy(), x["vel"].copy(), x["pos"].copy(), x["rot"].copy()))
        predict(acc, gyro, t1 - cache[-1][2])
        self.last_imu = cache[-1]
        self.last_end = t1
        R, p, Re, pe = x["rot"], x["pos"], x["rot_ext"], x["pos_ext"]
        out = cloud.copy()
        i = len(cloud) - 1
        for k in range(len(poses) - 1, 0, -1):
            h_off, _, _, h_vel, h_pos, h_rot = poses[k - 1]
            _, t_acc, t_gyro = poses[k][:3]
            # the points are edited IN PLACE and the iterator stays on the first point once it gets there: a first point later than the head of
            # an earlier pose pair is compensated again, from its already compensated coordinates (the reference's loop, literally)
            while float(out[i, 3]) / 1000.0 > h_off:
                dt = float(out[i, 3]) / 1000.0 - h_off
                pt = out[i, :3].astype(np.float64)
                point_rot = h_rot @ _so3_exp(t_gyro * dt)
                point_pos = h_pos + h_vel * dt + 0.5 * t_acc * dt * dt
                out[i, :3] = (Re.T @ (R.T @ (point_rot @ (Re @ pt + pe) + point_pos - p) - pe)).astype(np.float32)
                if i == 0:
                    break
                i -= 1
        return out


def test_imu_init_and_motion_compensation_match_numpy_restatement(oracle_mod):
    """The oracle's initializeImu / undistortCloud (and with them the priors it feeds the update) against the numpy restatement above, scan by
    scan along a moving synthetic sequence, each scan started from the oracle's own posterior: propagated prior to 1e-11, compensated
    points equal up to one float32 rounding (the result is stored as float32 on both sides)."""
    from voxelmapplus_fastlio2_b200 import synth
    cfg = default_config(max_points_per_scan=4096)
    o = oracle_mod.Oracle(cfg)
    ref = _LioNumpy(cfg)
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=1500))
    checked = moved = 0
    x = None
    for pk in seq.packages(34):
        raw = pk.cloud.copy()
        assert np.all(np.diff(raw[:, 3]) >= 0)                      # already in time order: the sort of lio_builder.cpp:75 is the identity
        _, _, status = o.lio_state()
        cloud = pk.cloud.copy()
        o.lio_process(pk.imus, cloud, pk.t0, pk.t1)
        xo, _, _ = o.lio_state()
        if status == 0:
            x = ref.initialize(pk.imus, pk.t1)
            if x is not None:
                d = xo.as_dict()
                np.testing.assert_allclose(d["rot"].reshape(3, 3), x["rot"], atol=1e-12)
                np.testing.assert_allclose(d["bg"], x["bg"], atol=1e-15)
                np.testing.assert_allclose(d["g"], x["g"], atol=1e-12)
            continue
        mine = ref.undistort(x, pk.imus, raw, pk.t0, pk.t1)
        x0, _ = o.get_prior() if status == 2 else (None, None)
        if x0 is not None:                                           # the prior of the update = the propagated state
            d0 = x0.as_dict()
            np.testing.assert_allclose(d0["pos"], x["pos"], atol=1e-11); np.testing.assert_allclose(d0["vel"], x["vel"], atol=1e-11)
            np.testing.assert_allclose(d0["rot"].reshape(3, 3), x["rot"], atol=1e-11)
        scale = float(np.abs(mine[:, :3]).max())
        dev = np.abs(mine[:, :3].astype(np.float64) - cloud[:, :3].astype(np.float64)).max()
        assert dev <= 2 * float(np.finfo(np.float32).eps) * scale, (pk.index, dev)
        assert (mine[:, :3] == cloud[:, :3]).mean() > 0.99
        moved = max(moved, float(np.abs(mine[:, :3] - raw[:, :3]).max()))
        checked += 1
        # next scan starts from the oracle's posterior (the numpy side restates the propagation, not the update)
        d = xo.as_dict()
        x = dict(pos=d["pos"], vel=d["vel"], ba=d["ba"], bg=d["bg"], g=d["g"], rot=d["rot"].reshape(3, 3), rot_ext=d["rot_ext"].reshape(3, 3), pos_ext=d["pos_ext"])
    assert checked >= 25 and moved > 0.01, (checked, moved)         # the compensation really moves points (centimetres)


# --------------------------------------------------------------------------- lidarToWorld + pv_list (what the map update is fed)
def test_world_points_and_pv_cov_match_numpy_restatement(oracle_mod):
    """The tail of LIOBuilder::process (lio_builder.cpp:155-163, 231-245): the float32 world transform and pv.cov with the POSTERIOR pose and
    covariance, restated with numpy.  The transform is evaluated in float32 with the association of pcl::transformPointCloud's SSE path
    (x' = m0 x + (m1 y + (m2 z + t)), the one assumption about PCL the oracle makes - stated here, not verified: PCL is not in the image); with it the
    world points are identical bit for bit; pv.cov to 1e-12 relative."""
    from voxelmapplus_fastlio2_b200 import synth
    cfg = default_config(max_points_per_scan=4096)
    o = oracle_mod.Oracle(cfg)
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=1200))
    checked = 0
    for pk in seq.packages(30):
        cloud = pk.cloud.copy()
        st = o.lio_process(pk.imus, cloud, pk.t0, pk.t1)
        x, P, status = o.lio_state()
        if status < 2 or st.iters == 0 or pk.index < 22:
            continue
        d = x.as_dict()
        R, Re = d["rot"].reshape(3, 3), d["rot_ext"].reshape(3, 3)
        r_wl, p_wl = R @ Re, R @ d["pos_ext"] + d["pos"]
        m, t = r_wl.astype(np.float32), p_wl.astype(np.float32)
        px, py, pz = cloud[:, 0], cloud[:, 1], cloud[:, 2]                   # the compensated cloud (float32), scan_resolution = 0
        w = np.stack([m[r, 0] * px + (m[r, 1] * py + (m[r, 2] * pz + t[r])) for r in range(3)], 1)
        assert w.dtype == np.float32
        pw, cw = o.dump_world_points(len(cloud))
        assert np.array_equal(pw, w.astype(np.float64)), f"scan {pk.index}: world points differ"
        worst = 0.0
        for i in range(0, len(cloud), 7):
            pl, cl = _body_cov_numpy(cloud[i, :3].astype(np.float64))
            ref = r_wl @ cl @ r_wl.T + _hat(pl) @ P[3:6, 3:6] @ _hat(pl).T + P[0:3, 0:3]
            worst = max(worst, np.abs(cw[i].reshape(3, 3) - ref).max() / np.abs(ref).max())
        assert worst < 1e-12, (pk.index, worst)
        checked += 1
    assert checked >= 5


# --------------------------------------------------------------------------- everything together, free-running
import pytest


@pytest.mark.parametrize("estimate_ext,scan_resolution,capacity", [(0, 0.0, 100000), (1, 0.0, 100000), (0, 0.1, 100000), (0, 0.0, 500)])
def test_free_running_python_lio_tracks_the_oracle(oracle_mod, estimate_ext, scan_resolution, capacity):
    """tests/lio_pyref.py - the pieces above composed into a second, free-running implementation of LIOBuilder::process (numpy / scipy / LAPACK,
    dict + OrderedDict map, persistent residual records) - beside the C++ oracle from the first IMU sample on, neither side ever seeing the other's
    state: iteration counts and effect_num per iteration identical, position within 1e-10 m, rotation matrix within 1e-10, posterior covariance 1e-8
    relative (measured: 2e-13 / 1.5e-13 / 1e-12 over 30 updates), and at the end the same voxels with the same flags and counts."""
    from lio_pyref import LioPy
    from voxelmapplus_fastlio2_b200 import synth
    cfg = default_config(max_points_per_scan=2048, map_capacity=capacity, estimate_ext=estimate_ext, scan_resolution=scan_resolution)
    o = oracle_mod.Oracle(cfg)
    py = LioPy(cfg)
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=1500 if scan_resolution > 0 else 700))    # (0.1 is the reference's default filter: pcl::VoxelGrid in the loop)
    updates = evictions = 0
    worst = dict(pos=0.0, rot=0.0, P=0.0)
    for pk in seq.packages(32):
        st = o.lio_process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        py.process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        evictions += len(py.map.evicted)
        xo, Po, status = o.lio_state()
        assert status == py.status
        if status < 2 or st.iters == 0:
            continue
        assert st.iters == py.iters and list(st.effect_num[:st.iters]) == py.effect, (pk.index, st.iters, py.iters, list(st.effect_num[:st.iters]), py.effect)
        d = xo.as_dict()
        worst["pos"] = max(worst["pos"], np.abs(d["pos"] - py.x["pos"]).max())
        worst["rot"] = max(worst["rot"], np.abs(d["rot"].reshape(3, 3) - py.x["rot"]).max())
        worst["P"] = max(worst["P"], np.abs(Po - py.P).max() / np.abs(Po).max())
        updates += 1
    print(f"python LIO vs oracle over {updates} updates: {worst}")
    assert updates >= 20 and (capacity > 500 or evictions > 500), (updates, evictions)     # the small map evicts continuously (LRU order is part of the state)
    assert worst["pos"] < 1e-10 and worst["rot"] < 1e-10 and worst["P"] < 1e-8, worst
    dm = o.dump_map()
    keys = [tuple(k) for k in dm["key"].tolist()]
    assert set(keys) == set(py.map.feat.keys())
    if capacity <= 500:
        assert keys == list(py.map.cache.keys())[::-1]              # the same LRU order, front first
    same = sum(int(v["n"]) == py.map.feat[k].n and bool(v["flags"] & F_PLANE) == py.map.feat[k].is_plane for v, k in zip(dm, keys))
    assert same >= 0.995 * len(keys), (same, len(keys))             # (a float32 world coordinate may round differently in the 9th digit of the pose)
