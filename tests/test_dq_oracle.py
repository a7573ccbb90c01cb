"""CPU tests of oracle/dq_oracle.c — the dual-quaternion kinematics downstream of the planner.

PARITY UNPINNED: dqrobotics (the library the reference calls, src/costp_controller.cpp:111-126, src/franka_robot.cpp:
6-22) is neither vendored nor installed, and the reference has no golden vectors for it. These tests anchor the
restatement on first principles instead: 4x4 homogeneous transforms of the modified Denavit-Hartenberg convention,
finite differences, and the Panda's published zero-configuration flange position."""
import numpy as np
import pytest

D = np.array([0.333, 0.0, 0.316, 0.0, 0.384, 0.0, 0.2104])          # src/franka_robot.cpp:9-10
A = np.array([0.0, 0.0, 0.0, 0.0825, -0.0825, 0.0, 0.088])            # :11
ALPHA = np.array([0.0, -1, 1, 1, -1, 1, 1]) * np.pi / 2                # :12


@pytest.fixture(scope="module")
def dq(oracle_built):
    return oracle_built.DqOracle()


def base_dq(p, quat_wxyz):
    """r + eps/2 p r (src/franka_robot.cpp:14-20)."""
    r = np.array(quat_wxyz, dtype=np.float64)
    r = r / np.linalg.norm(r)
    pq = np.array([0.0, *p])
    return np.concatenate([r, 0.5 * qmul(pq, r)])


def qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                     a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                     a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def rot_of(x):
    w, a, b, c = x[:4]
    return np.array([[1 - 2 * (b * b + c * c), 2 * (a * b - c * w), 2 * (a * c + b * w)],
                     [2 * (a * b + c * w), 1 - 2 * (a * a + c * c), 2 * (b * c - a * w)],
                     [2 * (a * c - b * w), 2 * (b * c + a * w), 1 - 2 * (a * a + b * b)]])


def homogeneous_fk(base_T, q):
    T = base_T.copy()
    for i in range(7):
        ca, sa, ct, st = np.cos(ALPHA[i]), np.sin(ALPHA[i]), np.cos(q[i]), np.sin(q[i])
        Rx = np.array([[1, 0, 0, A[i]], [0, ca, -sa, 0], [0, sa, ca, 0], [0, 0, 0, 1.0]])   # Rot_x(alpha) Trans_x(a)
        Rz = np.array([[ct, -st, 0, 0], [st, ct, 0, 0], [0, 0, 1, D[i]], [0, 0, 0, 1.0]])   # Rot_z(theta) Trans_z(d)
        T = T @ Rx @ Rz
    return T


IDENT = np.array([1.0, 0, 0, 0, 0, 0, 0, 0])


def test_zero_configuration_known_answer(dq):
    """Panda at q = 0: flange at x = 0.088, z = 0.333 + 0.316 + 0.384 - 0.107 = 0.926 (Franka's published DH); the
    reference's last link adds the 0.1034 m hand (d7 = 0.2104): z = 0.8226, tool axis pointing down."""
    x = dq.fkm(IDENT, np.zeros(7))
    np.testing.assert_allclose(dq.translation(x), [0.088, 0.0, 1.033 - 0.2104], atol=1e-15)
    np.testing.assert_allclose(rot_of(x) @ [0, 0, 1], [0, 0, -1], atol=1e-15)
    assert abs(np.dot(x[:4], x[:4]) - 1) < 1e-15 and abs(np.dot(x[:4], x[4:])) < 1e-15  # unit dual quaternion


@pytest.mark.parametrize("seed", range(6))
def test_fkm_equals_homogeneous_modified_dh(dq, seed):
    rng = np.random.default_rng(seed)
    q = rng.uniform(-2.5, 2.5, 7)
    p, quat = rng.uniform(-1, 1, 3), rng.normal(size=4)
    b = base_dq(p, quat)
    T0 = np.eye(4)
    T0[:3, :3], T0[:3, 3] = rot_of(b), p
    T = homogeneous_fk(T0, q)
    x = dq.fkm(b, q)
    np.testing.assert_allclose(rot_of(x), T[:3, :3], atol=2e-14)
    np.testing.assert_allclose(dq.translation(x), T[:3, 3], atol=2e-14)


@pytest.mark.parametrize("seed", range(4))
def test_jacobians_equal_finite_differences(dq, seed):
    rng = np.random.default_rng(100 + seed)
    q = rng.uniform(-2.0, 2.0, 7)
    b = base_dq(rng.uniform(-1, 1, 3), rng.normal(size=4))
    J, G = dq.pose_jacobian(b, q), dq.geom_jacobian(b, q)
    h = 1e-6
    for i in range(7):
        e = np.zeros(7)
        e[i] = h
        xp, xm = dq.fkm(b, q + e), dq.fkm(b, q - e)
        np.testing.assert_allclose(J[:, i], (xp - xm) / (2 * h), atol=1e-8)
        np.testing.assert_allclose(G[3:, i], (dq.translation(xp) - dq.translation(xm)) / (2 * h), atol=1e-8)
        W = (rot_of(xp) - rot_of(xm)) / (2 * h) @ rot_of(dq.fkm(b, q)).T  # skew(omega), base frame
        np.testing.assert_allclose(G[:3, i], [W[2, 1], W[0, 2], W[1, 0]], atol=1e-8)


def test_path_score_tracks_a_reachable_line_and_flags_an_unreachable_one(dq):
    lo = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])  # src/costp_controller.cpp:41-44
    hi = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])
    q0 = np.array([0.0, -0.4, 0.0, -2.0, 0.0, 1.6, 0.8])
    start = dq.translation(dq.fkm(IDENT, q0))
    line = start + np.linspace(0, 1, 150)[:, None] * np.array([0.2, 0.1, -0.05])  # ~1.5 mm per point
    s = dq.score_path(IDENT, q0, line, lo, hi, damping=1e-3, tol_pos=1e-3)
    assert s["feasible"] == 1 and s["first_bad_point"] == -1
    assert s["max_pos_err"] < 1e-3 and s["min_joint_margin"] > 0 and s["min_manipulability"] > 0.01
    np.testing.assert_allclose(dq.translation(dq.fkm(IDENT, s["q_final"])), line[-1], atol=1e-3)
    far = start + np.linspace(0, 1, 150)[:, None] * np.array([1.5, 0.0, 0.0])  # leaves the workspace (reach 0.855 m)
    s = dq.score_path(IDENT, q0, far, lo, hi, damping=1e-3, tol_pos=1e-3)
    assert s["feasible"] == 0 and 0 < s["first_bad_point"] < 150
