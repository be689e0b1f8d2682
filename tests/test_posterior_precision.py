"""How well is the posterior covariance of IESKF::update (ieskf.cpp:141-155) pinned, and how close is the CUDA path to it?

The reference forms H_ = J^T P^-1 J + [H 0; 0 0] with two 23x23 inverses (P, then H_) and P_post = L H_^-1 L^T.  The CUDA path
evaluates the algebraically identical matrix-inversion-lemma form (vmp_solve.cuh: one D x D inverse, no P^-1).  These tests compute
the exact posterior of the same inputs in 50-digit arithmetic (mpmath) and compare, entry by entry relative to sqrt(P_ii P_jj):
  * (CPU) the oracle's = the reference's fp64 evaluation against the exact value;
  * (CPU) the lemma form in plain fp64 (numpy) against the exact value;
  * (GPU) the device's P_post against the exact value and against the oracle's.
Round 1 compared the device's P with the oracle's at rtol 2e-5 and called that "the conditioning of either form"; the 50-digit
evaluation shows that both forms are good to ~1e-12 on these inputs, so the device is now held to 1e-9 like every tier-2 quantity.
"""
import numpy as np
import pytest

from test_oracle_math import _boxminus, _fields, _jac

mp = pytest.importorskip("mpmath")


def rel_err(a, b):
    """max |a_ij - b_ij| / sqrt(b_ii b_jj): the error of a covariance entry relative to the scale of its row and column"""
    sc = np.sqrt(np.abs(np.outer(np.diag(b), np.diag(b))))
    return float(np.max(np.abs(a - b) / sc))


def exact_posterior(x0, P0, xk, H12, b12, x_post):
    """ieskf.cpp:134-155 for the last executed iteration with every inverse / product in 50-digit arithmetic.
    Inputs are the fp64 quantities the reference holds at that point (state, prior, measurement H / b)."""
    mp.mp.dps = 50
    x_pred = _fields(x0)
    delta = _boxminus(xk, x_pred)
    J = mp.matrix(_jac(xk, x_pred, delta).tolist())
    Pinv = mp.inverse(mp.matrix(P0.tolist()))
    H_ = J.T * Pinv * J
    b_ = J.T * Pinv * mp.matrix(delta.tolist())
    for i in range(12):
        b_[i] += mp.mpf(float(b12[i]))
        for j in range(12):
            H_[i, j] += mp.mpf(float(H12[i, j]))
    Hinv = mp.inverse(H_)
    step = -(Hinv * b_)
    step64 = np.array([float(v) for v in step])
    L = mp.matrix(_jac(_fields(x_post), x_pred, step64).tolist())
    Pe = L * Hinv * L.T
    return np.array([[float(Pe[i, j]) for j in range(23)] for i in range(23)]), delta, step64


def lemma_posterior_fp64(x0, P0, xk, H12, b12, x_post, step64, D=6):
    """the device's algebra (vmp_solve.cuh) in plain numpy fp64: G = (I + P_DD S)^-1, Q = P[:, :D] S G, P_post = (L A)(P - Q P[:D, :])(L A)^T"""
    x_pred = _fields(x0)
    delta = _boxminus(xk, x_pred)
    A = np.linalg.inv(_jac(xk, x_pred, delta))          # block diagonal: exact 3x3 / 2x2 inverses
    S = A[:D, :D].T @ H12[:D, :D] @ A[:D, :D]
    G = np.linalg.inv(np.eye(D) + P0[:D, :D] @ S)
    Q = P0[:, :D] @ S @ G
    LA = _jac(_fields(x_post), x_pred, step64) @ A
    return LA @ (P0 - Q @ P0[:D, :]) @ LA.T


def _scans(om, n_scans=34, first=22):
    """oracle run; yields (package with the cloud compensated in place, oracle, stats, data of the last executed iteration or None)"""
    from voxelmapplus_fastlio2_b200 import synth
    from voxelmapplus_fastlio2_b200.ctypes_defs import default_config
    cfg = default_config(max_points_per_scan=4096)
    o = om.Oracle(cfg)
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=3000))
    for pk in seq.packages(n_scans):
        st = o.lio_process(pk.imus, pk.cloud, pk.t0, pk.t1)
        x_post, P_post, status = o.lio_state()
        data = None
        if status == 2 and st.iters >= 1 and pk.index >= first:          # moving by then: non-trivial deltas in every block
            x0, P0 = o.get_prior()
            k = st.iters - 1
            H12, b12 = o.get_iter_Hb(k)
            data = (x0, P0, _fields(o.get_iter_state(k)), H12, b12, x_post, P_post)
        yield pk, o, st, status, data


def test_reference_arithmetic_does_not_pin_P_to_1e9(oracle_mod):
    worst_o = worst_l = 0.0
    best_o = 1.0
    n = 0
    for pk, o, st, status, data in _scans(oracle_mod):
        if data is None:
            continue
        x0, P0, xk, H12, b12, x_post, P_post = data
        Pe, _, step64 = exact_posterior(x0, P0, xk, H12, b12, x_post)
        r_o = rel_err(P_post, Pe)
        r_l = rel_err(lemma_posterior_fp64(x0, P0, xk, H12, b12, x_post, step64), Pe)
        worst_o, worst_l, best_o = max(worst_o, r_o), max(worst_l, r_l), min(best_o, r_o)
        n += 1
    assert n >= 8
    print(f"posterior covariance vs 50-digit evaluation over {n} scans: reference form (oracle) {best_o:.1e} .. {worst_o:.1e}, lemma form (fp64) <= {worst_l:.1e}")
    assert worst_o < 1e-10, "the oracle's P is not pinned by the exact value"
    assert worst_l < 1e-10, "the lemma form loses accuracy on these inputs"


@pytest.mark.gpu
def test_device_posterior_is_as_close_to_exact_as_the_reference(oracle_mod):
    """vmp_scan from the oracle's prior on the oracle's compensated cloud: the device's P_post against the 50-digit value."""
    from voxelmapplus_fastlio2_b200.bindings import HotPath
    from voxelmapplus_fastlio2_b200.ctypes_defs import default_config
    g = HotPath(default_config(max_points_per_scan=4096))
    worst_o = worst_g = worst_go = 0.0
    n = 0
    for pk, o, st, status, data in _scans(oracle_mod):
        if status < 2:
            continue
        xyz = np.ascontiguousarray(pk.cloud[:, :3])
        x0, P0 = o.get_prior()
        if st.iters == 0:
            g.first_scan(x0, P0, xyz)
            continue
        xg, Pg, sg = g.scan(x0, P0, xyz)
        assert sg.iters == st.iters
        if data is None:
            continue
        _, _, xk, H12, b12, x_post, P_post = data
        Pe, _, _ = exact_posterior(x0, P0, xk, H12, b12, x_post)
        worst_o, worst_g = max(worst_o, rel_err(P_post, Pe)), max(worst_g, rel_err(Pg, Pe))
        worst_go = max(worst_go, rel_err(Pg, P_post))
        n += 1
    assert n >= 8
    print(f"posterior covariance over {n} scans: oracle vs exact <= {worst_o:.1e}, device vs exact <= {worst_g:.1e}, device vs oracle <= {worst_go:.1e}")
    assert worst_g <= 1e-9 and worst_go <= 1e-9, (worst_g, worst_go)
