"""Shared helpers of the parity tests: comparison of map dumps, synthetic map-update workloads."""
from __future__ import annotations

import numpy as np

from voxelmapplus_fastlio2_b200.ctypes_defs import F_INIT, F_MERGED, F_PLANE, F_UPDATE_ENABLE  # noqa: F401


def keyset(dump):
    return {tuple(k) for k in dump["key"].tolist()}


def sort_by_key(dump):
    k = dump["key"]
    order = np.lexsort((k[:, 2], k[:, 1], k[:, 0]))
    return dump[order]


def groups_partition(dump):
    """group ids are arbitrary numbers (Q22): compare the partition they induce, keyed by voxel key."""
    part = {}
    for k, g in zip(dump["key"].tolist(), dump["group"].tolist()):
        part.setdefault(g, []).append(tuple(k))
    return sorted(sorted(v) for v in part.values())


def assert_maps_equal(a, b, exact=True, rtol=1e-9, what=""):
    """a, b: HotPath.dump_map() arrays (LRU order).  Tier 1: keys, LRU order, flags, counts bit-exact;
    plane parameters bit-exact when `exact` (same inputs, same evaluation order) else within rtol."""
    assert len(a) == len(b), f"{what}: map sizes differ {len(a)} vs {len(b)}"
    assert np.array_equal(a["key"], b["key"]), f"{what}: keys / LRU order differ"
    for f in ("n", "n_temp", "newly_add_point", "flags"):
        if not np.array_equal(a[f], b[f]):
            bad = np.nonzero(a[f] != b[f])[0][:5]
            raise AssertionError(f"{what}: field {f} differs at {bad}: {a[f][bad]} vs {b[f][bad]} keys {a['key'][bad].tolist()}")
    assert groups_partition(a) == groups_partition(b), f"{what}: group partition differs"
    for f in ("mean", "ppt", "norm", "cov", "center"):
        if exact:
            # NaN == NaN counts as equal: degenerate voxels (all points identical -> lambda0 == lambda_m)
            # give inf/NaN covariances in the reference arithmetic as well, in the same places
            if not np.array_equal(a[f], b[f], equal_nan=True):
                d = np.abs(a[f] - b[f]).reshape(len(a), -1).max(axis=1)
                bad = np.argsort(-d)[:3]
                raise AssertionError(f"{what}: field {f} not bit-exact, worst abs diff {d[bad]} at keys {a['key'][bad].tolist()} "
                                     f"flags {a['flags'][bad]} n {a['n'][bad]}")
        else:
            # Q15: world points are float32.  A posterior that differs from the oracle's in the 9th digit flips the
            # float32 rounding of some coordinates by one ulp (2.4e-7 at 4 m; at 200 000 points per scan that is
            # every 7th voxel), so free-running maps are compared statistically; the bit-exact statement is made by
            # the tests that feed both sides the same world points.
            x, y = a[f].reshape(len(a), -1), b[f].reshape(len(b), -1)
            if len(x) == 0:
                continue
            eps32 = float(np.finfo(np.float32).eps)
            fin = np.where(np.isfinite(y), np.abs(y), 0.0).max(axis=1)
            if f in ("mean", "ppt", "center"):
                # moments: every voxel within rtol, or 4 float32 ulps of the field's largest magnitude
                ok = np.all(np.isclose(x, y, rtol=rtol, atol=4 * eps32 * float(fin.max()), equal_nan=True), axis=1)
                assert ok.all(), f"{what}: field {f} differs beyond float32-rounding effects at keys {a['key'][~ok][:5].tolist()}"
            else:
                # normals / covariances go through an eigen-decomposition whose conditioning is the eigenvalue gap:
                # 99 % of the voxels within 1e-4 of the voxel's own largest entry
                vmax = np.maximum(fin, 1e-300)[:, None]
                ok = np.all(np.isclose(x / vmax, y / vmax, rtol=0, atol=1e-4, equal_nan=True), axis=1)
                assert (~ok).sum() <= 0.01 * len(x), f"{what}: field {f}: {(~ok).sum()} of {len(x)} voxels differ by more than 1e-4 relative"


def cov_rel_err(a, b):
    """max |a_ij - b_ij| / sqrt(b_ii b_jj): the error of a covariance entry relative to the scale of its row and column
    (tests/test_posterior_precision.py pins both the oracle's and the device's posterior to ~1e-13 of the exact value this way)"""
    sc = np.sqrt(np.abs(np.outer(np.diag(b), np.diag(b))))
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / sc))


def plane_cloud(rng, n, origin, u, v, extent, noise):
    """n points on the rectangle origin + a u + b v, a,b in [0,extent), with gaussian noise along the normal."""
    u = np.asarray(u, float); v = np.asarray(v, float)
    nrm = np.cross(u, v); nrm /= np.linalg.norm(nrm)
    a = rng.uniform(0, extent[0], n); b = rng.uniform(0, extent[1], n)
    return np.asarray(origin, float) + a[:, None] * u + b[:, None] * v + rng.normal(0, noise, n)[:, None] * nrm


def random_cov(rng, n, scale=1e-3):
    """n mildly non-symmetric positive 3x3 matrices (pv.cov is not exactly symmetric in the reference either)."""
    A = rng.normal(0, 1, (n, 3, 3))
    S = np.einsum("nij,nkj->nik", A, A) * scale + np.eye(3) * scale
    S[:, 0, 1] *= (1 + 1e-12)
    return S.reshape(n, 9)


def wall_workload(seed, scans=12, pts=1500, voxel=0.5):
    """A few coplanar walls + clutter, revisited over several scans so that voxels fill, close and merge."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    out = []
    for s in range(scans):
        parts = [plane_cloud(rng, pts // 3, (0.1, 0.2, 0.0), (1, 0, 0), (0, 0, 1), (4.0, 2.0), 0.01),
                 plane_cloud(rng, pts // 3, (0.0, 0.3, 0.1), (0, 1, 0), (0, 0, 1), (3.0, 2.0), 0.01),
                 plane_cloud(rng, pts // 6, (0.2, 0.1, 0.02), (1, 0, 0), (0, 1, 0), (4.0, 3.0), 0.01),
                 rng.uniform(-1, 5, (pts - pts // 3 * 2 - pts // 6, 3))]
        p = np.concatenate(parts)
        p = p[rng.permutation(len(p))]
        # the reference inserts float32-rounded world points (Q15)
        p = p.astype(np.float32).astype(np.float64)
        out.append((p, random_cov(rng, len(p))))
    return out
