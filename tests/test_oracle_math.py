"""Oracle self-tests: the property tests g2o's own unit tests use plus an independent NumPy
restatement of the edge math, including the bug-compatible Jacobians (App. E)."""
import ctypes as C

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import oracle_lib
from nrslam_b200.abi import Camera

L = None


def setup_module(_m):
    global L
    oracle_lib.build()
    L = oracle_lib.lib()


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


CAMS = [Camera.pinhole(472.64955, 472.64955, 479.5, 359.5),
        Camera.kb8(717.2104, 717.4816, 735.3566, 552.7982, -0.1389272, -0.001239606, 0.0009125824, -4.071615e-05)]


def edge_eval(cam, etype, pose7, pts, vals):
    err = np.zeros(3)
    J = np.zeros(72)
    pts = np.ascontiguousarray(pts, np.float64).reshape(-1)
    L.orc_edge_eval(C.byref(cam), etype, dp(np.ascontiguousarray(pose7, np.float64)), dp(pts),
                    dp(np.ascontiguousarray(vals, np.float64)), dp(err), dp(J))
    return err, J.reshape(4, 18)


def np_project(cam, X):
    p = np.array(cam.params[:], np.float64)
    if cam.model == 0:
        return np.array([p[0] * X[0] / X[2] + p[2], p[1] * X[1] / X[2] + p[3]])
    r = np.hypot(X[0], X[1])
    th = np.arctan2(r, X[2])
    psi = np.arctan2(X[1], X[0])
    rd = th + p[4] * th ** 3 + p[5] * th ** 5 + p[6] * th ** 7 + p[7] * th ** 9
    return np.array([p[0] * rd * np.cos(psi) + p[2], p[1] * rd * np.sin(psi) + p[3]])


def test_huber_derivative_matches_central_difference():
    # third_party/g2o/unit_test/general/robust_kernel_tests.cpp:96-123 (delta 1e-9 there; 1e-6 is better
    # conditioned in double and the 1e-5 tolerance is g2o's)
    rho = np.zeros(3)
    for delta in (1.0, np.sqrt(5.99), np.sqrt(0.584)):
        for e in (0.1, 0.5, 2.0, 5.99, 6.5, 40.0, 1e3):
            L.orc_huber(C.c_double(e), C.c_double(delta), dp(rho))
            r0 = rho.copy()
            h = 1e-6 * max(e, 1.0)
            L.orc_huber(C.c_double(e + h), C.c_double(delta), dp(rho)); up = rho[0]
            L.orc_huber(C.c_double(e - h), C.c_double(delta), dp(rho)); dn = rho[0]
            if abs(e - delta * delta) > 2 * h:
                assert abs((up - dn) / (2 * h) - r0[1]) < 1e-5
            assert r0[0] <= e + 1e-12


@pytest.mark.parametrize("cam", CAMS)
def test_projection_and_jacobian_against_numpy(cam):
    rng = np.random.default_rng(0)
    for _ in range(50):
        X = np.array([rng.uniform(-1.5, 1.5), rng.uniform(-1.0, 1.0), rng.uniform(1.5, 4.0)])
        uv = np.zeros(2)
        J = np.zeros(6)
        L.orc_project(C.byref(cam), dp(X), dp(uv), dp(J))
        ref = np_project(cam, X)
        assert np.allclose(uv, ref, rtol=0, atol=2e-3)  # fp32 evaluation (camera_model.h:89-95): ~1e-4 px noise
        Jn = np.zeros((2, 3))
        for k in range(3):
            h = 1e-6
            d = np.zeros(3); d[k] = h
            Jn[:, k] = (np_project(cam, X + d) - np_project(cam, X - d)) / (2 * h)
        assert np.allclose(J.reshape(2, 3), Jn, rtol=2e-4, atol=2e-3)


def test_se3_exp_matches_scipy():
    rng = np.random.default_rng(1)
    for _ in range(20):
        u = rng.normal(size=6) * 0.3
        T = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)
        out = np.zeros(7)
        L.orc_se3_exp_mul(dp(u), dp(T), dp(out))
        q = Rotation.from_rotvec(u[:3]).as_quat()
        if q[3] < 0:
            q = -q
        assert np.allclose(out[:4], q, atol=1e-12)
        th = np.linalg.norm(u[:3])
        K = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * K + (th - np.sin(th)) / th ** 3 * K @ K
        assert np.allclose(out[4:], V @ u[3:], atol=1e-12)
        # left-multiplicative update: exp(u) * T
        T2 = np.concatenate([Rotation.from_rotvec([0.1, -0.2, 0.05]).as_quat(), [0.3, -0.1, 0.2]])
        L.orc_se3_exp_mul(dp(u), dp(T2), dp(out))
        Rn = Rotation.from_rotvec(u[:3]) * Rotation.from_quat(T2[:4])
        qn = Rn.as_quat()
        if qn[3] < 0:
            qn = -qn
        assert np.allclose(out[:4], qn, atol=1e-12)
        assert np.allclose(out[4:], Rotation.from_rotvec(u[:3]).apply(T2[4:]) + V @ u[3:], atol=1e-12)


@pytest.mark.parametrize("cam", CAMS)
def test_reprojection_edges_against_numpy(cam):
    """r = z - pi(R (X + d) + t); J_pose = -J_pi [ -[p]x | I ]; J_pt = -J_pi R  (SURVEY App. A)."""
    rng = np.random.default_rng(2)
    for etype in (0, 1, 2):
        for _ in range(10):
            q = Rotation.from_rotvec(rng.normal(size=3) * 0.05).as_quat()
            if q[3] < 0:
                q = -q
            t = rng.normal(size=3) * 0.05
            pose = np.concatenate([q, t])
            Xw = np.array([rng.uniform(-1, 1), rng.uniform(-0.8, 0.8), rng.uniform(2.5, 3.5)])
            d = rng.normal(size=3) * 0.02
            z = rng.uniform(100, 600, size=2)
            pts = np.zeros((4, 3))
            vals = np.zeros(14)
            vals[:2] = z
            if etype == 0:
                vals[3:6] = Xw; X = Xw
            elif etype == 1:
                vals[3:6] = Xw; pts[0] = d; X = Xw + d
            else:
                pts[0] = Xw; X = Xw
            err, J = edge_eval(cam, etype, pose, pts, vals)
            Rm = Rotation.from_quat(q).as_matrix()
            pc = Rm @ X + t
            assert np.allclose(err[:2], z - np_project(cam, pc), atol=3e-3)
            Jpi = np.zeros((2, 3))
            for k in range(3):
                h = 1e-6
                dd = np.zeros(3); dd[k] = h
                Jpi[:, k] = (np_project(cam, pc + dd) - np_project(cam, pc - dd)) / (2 * h)
            x, y, zz = pc
            M = np.array([[0, zz, -y, 1, 0, 0], [-zz, 0, x, 0, 1, 0], [y, -x, 0, 0, 0, 1]])
            assert np.allclose(J[0, :12].reshape(2, 6), -Jpi @ M, rtol=5e-4, atol=5e-2)
            if etype != 0:
                assert np.allclose(J[1, :6].reshape(2, 3), -Jpi @ Rm, rtol=5e-4, atol=5e-3)


def test_regulariser_edges_against_numpy_including_quirks():
    rng = np.random.default_rng(3)
    cam = CAMS[0]
    pose = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)
    for _ in range(10):
        pts = rng.normal(size=(4, 3)) * 0.3
        w, k, d0 = rng.uniform(0.3, 1.0), 1.1, rng.uniform(0.1, 0.4)
        r1, r2 = rng.normal(size=3), rng.normal(size=3)
        vals = np.zeros(14)
        vals[0] = d0; vals[6:9] = r1; vals[9:12] = r2; vals[12] = w; vals[13] = k
        # spatial with deformation: w (d_i - d_j), J = +-w I
        err, J = edge_eval(cam, 3, pose, pts, vals)
        assert np.allclose(err, w * (pts[0] - pts[1]))
        assert np.allclose(J[0, :9].reshape(3, 3), w * np.eye(3)) and np.allclose(J[1, :9].reshape(3, 3), -w * np.eye(3))
        # position with deformation: exact derivative of k (|c1 - c2| - d0) / d0
        err, J = edge_eval(cam, 4, pose, pts, vals)
        c = (r1 + pts[0]) - (r2 + pts[1])
        assert np.isclose(err[0], k * (np.linalg.norm(c) - d0) / d0)
        assert np.allclose(J[0, :3], k * c / (d0 * np.linalg.norm(c))) and np.allclose(J[1, :3], -J[0, :3])
        # spatial fixed: w (d - d_ref), J = w I
        err, J = edge_eval(cam, 5, pose, pts, vals)
        assert np.allclose(err, w * (pts[0] - pts[1])) and np.allclose(J[0, :9].reshape(3, 3), w * np.eye(3))
        # BA spring, quirk E1 (position_regularizer.cc:53-60): (k/d0) (1/sqrt|D|) 2 D — NOT the true derivative
        err, J = edge_eval(cam, 6, pose, pts, vals)
        D = pts[0] - pts[1]
        n = np.linalg.norm(D)
        assert np.isclose(err[0], k * (n - d0) / d0)
        assert np.allclose(J[0, :3], (k / d0) * 2 * D / np.sqrt(n)) and np.allclose(J[1, :3], -J[0, :3])
        true_J = k * D / (d0 * n)
        assert not np.allclose(J[0, :3], true_J)  # bug-compatible on purpose
        # damper: w ((x_i' - x_i) - (x_j' - x_j)), J = (-w, +w, +w, -w) I
        err, J = edge_eval(cam, 7, pose, pts, vals)
        assert np.allclose(err, w * ((pts[2] - pts[0]) - (pts[3] - pts[1])))
        for a, s in enumerate((-1, 1, 1, -1)):
            assert np.allclose(J[a, :9].reshape(3, 3), s * w * np.eye(3))
