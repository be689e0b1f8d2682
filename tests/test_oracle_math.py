"""Pins the oracle's restated third-party numerics (Eigen / Sophus, absent from this image) against
numpy / scipy: the oracle is only trustworthy as a checker if its small algebra is right."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from voxelmapplus_fastlio2_b200.ctypes_defs import VmpState


@pytest.fixture(scope="module")
def om(oracle_mod):
    return oracle_mod


def test_eig3_matches_numpy(om):
    rng = np.random.default_rng(0)
    for k in range(300):
        A = rng.normal(size=(3, 3)) * 10.0 ** rng.integers(-4, 3)
        S = A @ A.T
        if k % 7 == 0:      # nearly planar point sets: one tiny eigenvalue, the case the plane test decides on
            S = A[:, :2] @ A[:, :2].T + 1e-9 * np.eye(3)
        w, V = om.eig3(S)
        wn, _ = np.linalg.eigh(S)
        assert np.all(np.diff(w) >= 0), "eigenvalues must be ascending"
        np.testing.assert_allclose(w, wn, rtol=1e-10, atol=1e-12 * np.abs(wn).max())
        np.testing.assert_allclose(V.T @ V, np.eye(3), atol=1e-12)
        np.testing.assert_allclose(S @ V, V * w, atol=1e-10 * np.abs(wn).max())


def test_eig3_reads_lower_triangle_and_handles_diagonal(om):
    S = np.diag([3.0, 1.0, 2.0])
    w, V = om.eig3(S)
    np.testing.assert_array_equal(w, [1.0, 2.0, 3.0])
    S2 = S.copy(); S2[0, 1] = 123.0       # upper triangle must be ignored (Eigen reads the lower one)
    w2, _ = om.eig3(S2)
    np.testing.assert_array_equal(w2, w)
    w0, _ = om.eig3(np.zeros((3, 3)))
    np.testing.assert_array_equal(w0, [0, 0, 0])


def test_inverse23_matches_numpy(om):
    rng = np.random.default_rng(1)
    for _ in range(20):
        A = rng.normal(size=(23, 23))
        P = A @ A.T + 1e-3 * np.eye(23)
        inv = om.inverse23(P)
        np.testing.assert_allclose(inv @ P, np.eye(23), atol=1e-8)
        np.testing.assert_allclose(inv, np.linalg.inv(P), rtol=1e-7, atol=1e-9 * np.abs(np.linalg.inv(P)).max())
    # a matrix that needs pivoting
    A = rng.normal(size=(23, 23)); A[0, 0] = 0.0
    np.testing.assert_allclose(om.inverse23(A) @ A, np.eye(23), atol=1e-9)


def test_so3_exp_log_jacobian(om):
    rng = np.random.default_rng(2)
    for _ in range(200):
        w = rng.normal(size=3) * rng.choice([1e-12, 1e-6, 0.1, 1.0, 3.0])
        R = om.so3_exp(w)
        np.testing.assert_allclose(R, Rotation.from_rotvec(w).as_matrix(), atol=1e-13)
        if np.linalg.norm(w) < np.pi - 1e-3:
            np.testing.assert_allclose(om.so3_log(R), w, atol=1e-12)
    w = np.array([0.3, -0.2, 0.5])
    J = om.left_jacobian(w)
    eps = 1e-7
    Jn = np.zeros((3, 3))
    for k in range(3):
        d = np.zeros(3); d[k] = eps
        Jn[:, k] = Rotation.from_matrix(om.so3_exp(w + d) @ om.so3_exp(w).T).as_rotvec() / eps
    np.testing.assert_allclose(J, Jn, atol=1e-6)
    np.testing.assert_allclose(om.left_jacobian(np.zeros(3)), np.eye(3))


def test_boxplus_boxminus_roundtrip(om):
    rng = np.random.default_rng(3)
    x = VmpState.identity()
    x.rot[:] = Rotation.from_rotvec([0.1, 0.2, -0.3]).as_matrix().ravel()
    x.pos[:] = [1, 2, 3]
    for _ in range(50):
        d = rng.normal(size=23) * 1e-2
        y = om.boxplus(x, d)
        back = om.boxminus(y, x)
        np.testing.assert_allclose(back, d, atol=1e-9)
        assert abs(np.linalg.norm(y.g[:]) - 9.81) < 1e-9, "gravity stays on the sphere"
    np.testing.assert_array_equal(om.boxminus(x, x), np.zeros(23))


def test_calc_body_cov(om):
    """commons.cpp:18-45 in numpy, incl. the z == 0 edit (Q16) and PCL's truncated DEG2RAD."""
    rng = np.random.default_rng(4)
    for k in range(50):
        p = rng.normal(size=3) * 10
        if k == 0:
            p[2] = 0.0
        pe, cov = om.calc_body_cov(p, 0.04, 0.1)
        q = p.copy()
        if q[2] == 0:
            q[2] = 0.001
        np.testing.assert_array_equal(pe, q)
        r = np.linalg.norm(q); d = q / r
        dh = np.array([[0, -d[2], d[1]], [d[2], 0, -d[0]], [-d[1], d[0], 0]])
        b1 = np.array([1, 1, -(d[0] + d[1]) / d[2]]); b1 /= np.linalg.norm(b1)
        b2 = np.cross(b1, d); b2 /= np.linalg.norm(b2)
        N = np.stack([b1, b2], 1)
        A = r * dh @ N
        s2 = np.sin(0.1 * 0.017453293) ** 2
        ref = np.outer(d, d) * 0.04 ** 2 + A @ (np.eye(2) * s2) @ A.T
        np.testing.assert_allclose(cov, ref, rtol=1e-12, atol=1e-18)


def test_voxel_key_floor_semantics(om):
    from voxelmapplus_fastlio2_b200.ctypes_defs import default_config
    o = om.Oracle(default_config(voxel_size=0.5, max_points_per_scan=16))
    np.testing.assert_array_equal(o.voxel_key([0.49, -0.01, -0.5]), [0, -1, -1])
    np.testing.assert_array_equal(o.voxel_key([0.5, -0.5000001, 1e6]), [1, -2, 2000000])


# ---------------------------------------------------------------------------------------------------------------------
# IESKF::update (ieskf.cpp:125-156) restated once more, independently, in dense numpy with scipy rotations: the oracle's
# manifold operators, Jacobians, the 23x23 algebra and the posterior covariance against it on a real update.
G0 = 9.81


def _hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


def _exp(w):
    return Rotation.from_rotvec(w).as_matrix()


def _log(R):
    return Rotation.from_matrix(R).as_rotvec()


def _left_jac(w):
    th = np.linalg.norm(w)
    if th < 1e-10:
        return np.eye(3) + 0.5 * _hat(w)
    O = _hat(w)
    return np.eye(3) + (1 - np.cos(th)) / th ** 2 * O + (th - np.sin(th)) / th ** 3 * O @ O


def _right_jac(w):
    return _left_jac(w).T


def _Bx(g):
    r = np.array([[-g[1], -g[2]],
                  [G0 - g[1] * g[1] / (G0 + g[0]), -g[2] * g[1] / (G0 + g[0])],
                  [-g[2] * g[1] / (G0 + g[0]), G0 - g[2] * g[2] / (G0 + g[0])]])
    return r / G0


def _Mx_res(g, res):
    bx = _Bx(g)
    bu = bx @ res
    return -_exp(bu) @ _hat(g) @ _left_jac(bu).T @ bx


def _Nx(g):
    return 1 / G0 / G0 * _Bx(g).T @ _hat(g)


def _fields(x):
    d = x.as_dict()
    d["rot"] = d["rot"].reshape(3, 3)
    d["rot_ext"] = d["rot_ext"].reshape(3, 3)
    return d


def _boxminus(a, b):
    """a - b, ieskf.cpp:35-71 (the generic branch of the gravity chart)."""
    d = np.zeros(23)
    d[0:3] = a["pos"] - b["pos"]
    d[3:6] = _log(b["rot"].T @ a["rot"])
    d[6:9] = _log(b["rot_ext"].T @ a["rot_ext"])
    d[9:12] = a["pos_ext"] - b["pos_ext"]
    d[12:15] = a["vel"] - b["vel"]
    d[15:18] = a["bg"] - b["bg"]
    d[18:21] = a["ba"] - b["ba"]
    v_sin = np.linalg.norm(_hat(a["g"]) @ b["g"])
    v_cos = a["g"] @ b["g"]
    theta = np.arctan2(v_sin, v_cos)
    if v_sin < 1e-11:
        d[21:23] = [3.1415926, 0] if abs(theta) > 1e-11 else [0, 0]
    else:
        d[21:23] = theta / v_sin * _Bx(b["g"]).T @ _hat(b["g"]) @ a["g"]
    return d


def _boxplus(a, d):
    o = {k: v.copy() for k, v in a.items()}
    o["pos"] = a["pos"] + d[0:3]
    o["rot"] = a["rot"] @ _exp(d[3:6])
    o["rot_ext"] = a["rot_ext"] @ _exp(d[6:9])
    o["pos_ext"] = a["pos_ext"] + d[9:12]
    o["vel"] = a["vel"] + d[12:15]
    o["bg"] = a["bg"] + d[15:18]
    o["ba"] = a["ba"] + d[18:21]
    o["g"] = _exp(_Bx(a["g"]) @ d[21:23]) @ a["g"]
    return o


def _jac(x, x_pred, delta):
    J = np.eye(23)
    J[3:6, 3:6] = _right_jac(delta[3:6])
    J[6:9, 6:9] = _right_jac(delta[6:9])
    J[21:23, 21:23] = _Nx(x["g"]) @ _Mx_res(x_pred["g"], delta[21:23])
    return J


def test_ieskf_update_matches_dense_numpy_restatement(om):
    from voxelmapplus_fastlio2_b200 import synth
    from voxelmapplus_fastlio2_b200.ctypes_defs import default_config
    o = om.Oracle(default_config(max_points_per_scan=4096))
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=3000))
    checked = 0
    for pk in seq.packages(30):
        st = o.lio_process(pk.imus, pk.cloud.copy(), pk.t0, pk.t1)
        x_post, P_post, status = o.lio_state()
        if status < 2 or st.iters < 1 or pk.index < 22:          # moving by then: non-trivial deltas in every block
            continue
        x0, P0 = o.get_prior()
        x_pred = _fields(x0)
        Pinv = np.linalg.inv(P0)
        for k in range(st.iters):
            xk = _fields(o.get_iter_state(k))
            H12, b12 = o.get_iter_Hb(k)
            delta = _boxminus(xk, x_pred)
            J = _jac(xk, x_pred, delta)
            b_ = J.T @ Pinv @ delta
            H_ = J.T @ Pinv @ J
            H_[:12, :12] += H12
            b_[:12] += b12
            step = -np.linalg.solve(H_, b_)
            x_next = _boxplus(xk, step)
            ref = _fields(o.get_iter_state(k + 1)) if k + 1 < st.iters else _fields(x_post)
            for f in ("pos", "rot", "rot_ext", "pos_ext", "vel", "bg", "ba", "g"):
                np.testing.assert_allclose(x_next[f], ref[f], rtol=0, atol=2e-9, err_msg=f"scan {pk.index} iteration {k} field {f}")
            checked += 1
        # posterior covariance from the last executed iteration (Q6): L from the final step and the updated state
        L = _jac(_fields(x_post), x_pred, step)
        P_np = L @ np.linalg.inv(H_) @ L.T
        np.testing.assert_allclose(P_post, P_np, rtol=1e-6, atol=1e-14)
    assert checked >= 10


def test_ieskf_predict_matches_dense_numpy_restatement(om):
    """IESKF::predict (ieskf.cpp:101-123): state propagation and P = F P F^T + G Q G^T against dense numpy."""
    from voxelmapplus_fastlio2_b200.ctypes_defs import default_config
    cfg = default_config(max_points_per_scan=64, map_capacity=64)
    o = om.Oracle(cfg)
    rng = np.random.Generator(np.random.Philox(key=5))
    x = VmpState.identity()
    x.pos[:] = [0.3, -1.2, 0.8]
    x.rot[:] = Rotation.from_rotvec([0.2, -0.4, 0.9]).as_matrix().ravel()
    x.rot_ext[:] = Rotation.from_rotvec([0.01, 0.02, -0.015]).as_matrix().ravel()
    x.pos_ext[:] = [0.05, -0.02, 0.1]
    x.vel[:] = [0.7, 0.1, -0.2]
    x.bg[:] = [0.002, -0.001, 0.0015]
    x.ba[:] = [0.01, -0.02, 0.015]
    gdir = np.array([0.05, -0.03, -1.0]); x.g[:] = (gdir / np.linalg.norm(gdir) * G0)
    A = rng.normal(0, 1, (23, 23))
    P = A @ A.T * 1e-3 + np.eye(23) * 1e-4
    o.set_state(x, P)
    Q = np.zeros((12, 12))
    Q[0:3, 0:3] = np.eye(3) * cfg.ng; Q[3:6, 3:6] = np.eye(3) * cfg.na
    Q[6:9, 6:9] = np.eye(3) * cfg.nbg; Q[9:12, 9:12] = np.eye(3) * cfg.nba
    xs = _fields(x)
    for step in range(5):
        acc = np.array([0.3, -0.2, 9.7]) + rng.normal(0, 0.2, 3)
        gyro = np.array([0.1, -0.3, 0.25]) + rng.normal(0, 0.05, 3)
        dt = 0.005 * (1 + step % 2)
        w, a = gyro - xs["bg"], acc - xs["ba"]
        F = np.eye(23)
        F[0:3, 12:15] = np.eye(3) * dt
        F[3:6, 3:6] = _exp(-w * dt)
        F[3:6, 15:18] = -_right_jac(w * dt) * dt
        F[12:15, 3:6] = -xs["rot"] @ _hat(a) * dt
        F[12:15, 18:21] = -xs["rot"] * dt
        Mx = -_hat(xs["g"]) @ _Bx(xs["g"])
        F[12:15, 21:23] = Mx * dt
        F[21:23, 21:23] = _Nx(xs["g"]) @ Mx
        G = np.zeros((23, 12))
        G[3:6, 0:3] = -_right_jac(w * dt) * dt
        G[12:15, 3:6] = -xs["rot"] * dt
        G[15:18, 6:9] = np.eye(3) * dt
        G[18:21, 9:12] = np.eye(3) * dt
        nxt = {k: v.copy() for k, v in xs.items()}
        nxt["pos"] = xs["pos"] + xs["vel"] * dt
        nxt["rot"] = xs["rot"] @ _exp(w * dt)
        nxt["vel"] = xs["vel"] + (xs["rot"] @ a + xs["g"]) * dt
        P = F @ P @ F.T + G @ Q @ G.T
        o.predict(acc, gyro, dt)
        xo, Po = o.get_state()
        got = _fields(xo)
        for f in ("pos", "rot", "rot_ext", "pos_ext", "vel", "bg", "ba", "g"):
            np.testing.assert_allclose(got[f], nxt[f], rtol=0, atol=1e-12, err_msg=f"step {step} field {f}")
        np.testing.assert_allclose(Po, P, rtol=1e-10, atol=1e-16)
        xs = nxt
