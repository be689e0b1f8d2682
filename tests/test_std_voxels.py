"""SURVEY.md 8(f) row 4 — STDManager::buildVoxels (std_matcher/src/std_manager/descriptor.cpp:70-122).

CPU part: the oracle restatement against an independent numpy statement of the same pass (float64 accumulation in point order,
`numpy.linalg.eig` = LAPACK dgeev, the general real solver that Eigen::EigenSolver is an implementation of).
GPU part: the device pass through the C ABI against the oracle: keys / counts / flags / sum / ppt / mean bit-exact, eigenvalues to
1e-9 relative (tier 2), eigenvectors up to sign.
"""
import numpy as np
import pytest

from oracle import oracle_py
from voxelmapplus_fastlio2_b200.ctypes_defs import STD_F_PLANE, STD_F_VALID

LAMBDA_RTOL = 1e-9      # tier 2: eigenvalues, relative to the largest one of the voxel
VEC_ATOL = 1e-6         # eigenvectors up to sign, for voxels whose eigenvalues are separated (gap > 1e-3 of the largest)


def submap_cloud(seed=0, n=60000, extent=40.0):
    """Walls, ground and clutter of a street-like sub-map, float32 xyzi."""
    rng = np.random.default_rng(seed)
    k = n // 4
    ground = np.c_[rng.uniform(-extent, extent, k), rng.uniform(-extent, extent, k), rng.normal(0.0, 0.02, k)]
    wall_a = np.c_[rng.uniform(-extent, extent, k), np.full(k, 12.3) + rng.normal(0, 0.02, k), rng.uniform(0, 6, k)]
    wall_b = np.c_[np.full(k, -17.7) + rng.normal(0, 0.02, k), rng.uniform(-extent, extent, k), rng.uniform(0, 6, k)]
    clutter = rng.uniform(-extent, extent, (n - 3 * k, 3)) * [1, 1, 0.1]
    pts = np.vstack([ground, wall_a, wall_b, clutter])
    pts = pts[rng.permutation(len(pts))]
    return np.c_[pts, rng.uniform(0, 255, len(pts))].astype(np.float32)


def numpy_build_voxels(cloud, voxel_size, min_point, thresh):
    """Independent statement of descriptor.cpp:70-122."""
    pts = cloud[:, :3].astype(np.float64)
    keys = np.floor(pts / voxel_size).astype(np.int64)
    order, nodes = [], {}
    for i, k in enumerate(map(tuple, keys)):
        nd = nodes.get(k)
        if nd is None:
            nd = nodes[k] = {"sum": np.zeros(3), "ppt": np.zeros((3, 3)), "n": 0}
            order.append(k)
        p = pts[i]
        nd["sum"] = nd["sum"] + p
        nd["ppt"] = nd["ppt"] + np.outer(p, p)
        nd["n"] += 1
    out = []
    for k in order:
        nd = nodes[k]
        rec = dict(key=k, count=nd["n"], sum=nd["sum"], ppt=nd["ppt"], valid=nd["n"] > min_point, plane=False)
        if rec["valid"]:
            mean = nd["sum"] / float(nd["n"])
            cov = nd["ppt"] / float(nd["n"]) - np.outer(mean, mean)
            w, v = np.linalg.eig(cov)
            w, v = w.real, v.real
            imin, imax = int(np.argmin(w)), int(np.argmax(w))
            imid = 3 - imin - imax
            rec.update(mean=mean, cov=cov)
            if w[imin] < thresh:
                rec.update(plane=True, lamdas=np.array([w[imin], w[imid], w[imax]]), norms=np.c_[v[:, imin], v[:, imid], v[:, imax]])
        out.append(rec)
    return out


def check_against_numpy(got, want):
    assert len(got) == len(want)
    n_plane = 0
    for g, w in zip(got, want):
        assert tuple(g["key"]) == w["key"] and g["count"] == w["count"]
        assert np.array_equal(g["sum"], w["sum"]) and np.array_equal(g["ppt"].reshape(3, 3), w["ppt"])      # same order, same bits
        assert bool(g["flags"] & STD_F_VALID) == w["valid"]
        if not w["valid"]:
            continue
        assert np.array_equal(g["mean"], w["mean"])
        if not w["plane"] and not (g["flags"] & STD_F_PLANE):
            continue
        if bool(g["flags"] & STD_F_PLANE) != w["plane"]:
            # only legitimate when lambda_min sits on the threshold to rounding
            lam_min = np.linalg.eigvalsh(w["cov"])[0]
            assert abs(lam_min - 0.01) < 1e-12, (w["key"], lam_min)
            continue
        n_plane += 1
        scale = max(abs(w["lamdas"][2]), 1e-300)
        assert np.max(np.abs(g["lamdas"] - w["lamdas"])) <= LAMBDA_RTOL * scale
        gaps = np.diff(w["lamdas"]) / scale
        N = g["norms"].reshape(3, 3)                   # row k = eigenvector k
        for k in range(3):
            assert abs(np.linalg.norm(N[k]) - 1.0) < 1e-12
            sep = min(gaps[k - 1] if k > 0 else 1.0, gaps[k] if k < 2 else 1.0)
            if sep > 1e-3:
                assert min(np.max(np.abs(N[k] - w["norms"][:, k])), np.max(np.abs(N[k] + w["norms"][:, k]))) < VEC_ATOL
    return n_plane


def test_oracle_build_voxels_against_numpy_lapack():
    cloud = submap_cloud(seed=3, n=20000)
    got = oracle_py.std_build_voxels(cloud, 1.0, 10, 0.01)
    want = numpy_build_voxels(cloud, 1.0, 10, 0.01)
    n_plane = check_against_numpy(got, want)
    assert n_plane > 100 and sum(1 for w in want if w["valid"] and not w["plane"]) > 10 and sum(1 for w in want if not w["valid"]) > 100


def test_oracle_build_voxels_edge_cases():
    assert len(oracle_py.std_build_voxels(np.zeros((0, 4), np.float32))) == 0
    # exactly voxel_min_point points: `size() <= voxel_min_point` skips it (descriptor.cpp:93); one more makes it valid
    rng = np.random.default_rng(0)
    p = np.c_[rng.uniform(0.1, 0.9, (11, 2)), np.full(11, 0.5), np.zeros(11)].astype(np.float32)
    v = oracle_py.std_build_voxels(p[:10], 1.0, 10, 0.01)
    assert len(v) == 1 and v[0]["count"] == 10 and v[0]["flags"] == 0 and not v[0]["mean"].any()
    v = oracle_py.std_build_voxels(p, 1.0, 10, 0.01)
    assert v[0]["flags"] == STD_F_VALID | STD_F_PLANE and abs(abs(v[0]["norms"][2]) - 1.0) < 1e-12      # normal = z
    # negative coordinates floor towards -inf (VoxelKey::index, descriptor.cpp:42-47)
    v = oracle_py.std_build_voxels(np.array([[-0.25, 0.25, -1.0, 0]], np.float32), 0.5, 10, 0.01)
    assert tuple(v[0]["key"]) == (-1, 0, -2)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,vs,minpt,thr", [(1, 60000, 1.0, 10, 0.01), (2, 200000, 0.5, 5, 0.005), (5, 3000, 2.0, 10, 0.01)])
def test_device_build_voxels_matches_oracle(seed, n, vs, minpt, thr):
    from voxelmapplus_fastlio2_b200.bindings import std_build_voxels
    cloud = submap_cloud(seed=seed, n=n)
    dev = std_build_voxels(cloud, vs, minpt, thr)
    orc = oracle_py.std_build_voxels(cloud, vs, minpt, thr)
    assert len(dev) == len(orc) > 50
    for f in ("key", "count", "flags", "sum", "ppt", "mean"):                   # tier 1
        assert np.array_equal(dev[f], orc[f]), f
    pl = (orc["flags"] & STD_F_PLANE) != 0
    assert pl.sum() > (50 if n >= 60000 else 0)
    scale = np.abs(orc["lamdas"][pl, 2:3])
    assert np.max(np.abs(dev["lamdas"][pl] - orc["lamdas"][pl]) / scale) <= LAMBDA_RTOL                     # tier 2
    # the device and the oracle run the same symmetric solver in the same order of operations: in practice identical
    d = np.minimum(np.abs(dev["norms"] - orc["norms"]).max(axis=1), np.abs(dev["norms"] + orc["norms"]).max(axis=1))
    assert d.max() < 1e-9
    assert not dev["lamdas"][~pl].any() and not dev["norms"][~pl].any()


@pytest.mark.gpu
def test_device_build_voxels_edges_and_errors():
    from voxelmapplus_fastlio2_b200.bindings import VmpError, std_build_voxels
    assert len(std_build_voxels(np.zeros((0, 4), np.float32))) == 0
    one = std_build_voxels(np.array([[-0.25, 0.25, -1.0, 0]], np.float32), 0.5, 10, 0.01)
    assert len(one) == 1 and tuple(one[0]["key"]) == (-1, 0, -2) and one[0]["count"] == 1 and one[0]["flags"] == 0
    # every point in one voxel: the longest ordered accumulation
    rng = np.random.default_rng(4)
    blob = np.c_[rng.uniform(0.01, 0.99, (50000, 2)), rng.normal(0.5, 0.01, 50000), np.zeros(50000)].astype(np.float32)
    dev, orc = std_build_voxels(blob), oracle_py.std_build_voxels(blob)
    assert len(dev) == 1 and dev[0]["count"] == 50000
    for f in ("sum", "ppt", "mean", "flags"):
        assert np.array_equal(dev[f], orc[f]), f
    with pytest.raises(VmpError):
        std_build_voxels(np.array([[np.nan, 0, 0, 0]], np.float32))
    with pytest.raises(VmpError):
        std_build_voxels(blob, voxel_size=0.0)
