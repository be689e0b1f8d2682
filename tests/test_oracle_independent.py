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
