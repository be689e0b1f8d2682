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
