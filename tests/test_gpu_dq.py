"""GPU tests of the downstream kinematics: the device dual-quaternion algebra and the batched
path-feasibility score against oracle/dq_oracle.c. Tolerances, not bit equality: the two sides call different
sin / cos implementations (CUDA's and glibc's differ in the last bits), everything else is the same IEEE
arithmetic; kinematics agree to 1e-13 absolute, scores (hundreds of dependent least-squares steps) to 1e-9 relative.
Parity with a dqrobotics build is unpinned (see the oracle's header)."""
import numpy as np
import pytest

from pmaf_b200 import loop, scenarios

pytestmark = pytest.mark.gpu

IDENT = np.array([1.0, 0, 0, 0, 0, 0, 0, 0])


def _planner():
    from pmaf_b200.planner import CfManager

    return CfManager(0)


@pytest.fixture(scope="module")
def dq(oracle_built):
    return oracle_built.DqOracle()


@pytest.mark.parametrize("seed", range(5))
def test_device_kinematics_match_oracle(dq, seed):
    rng = np.random.default_rng(seed)
    q = rng.uniform(-2.5, 2.5, 7)
    r = rng.normal(size=4)
    r /= np.linalg.norm(r)
    p = rng.uniform(-1, 1, 3)
    pq = np.array([0.0, *p])
    qm = lambda a, b: np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                                a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])
    base = np.concatenate([r, 0.5 * qm(pq, r)])
    m = _planner()
    pose, J, G = m.dq_kinematics(base, q)
    m.close()
    np.testing.assert_allclose(pose, dq.fkm(base, q), atol=1e-13)
    np.testing.assert_allclose(J, dq.pose_jacobian(base, q), atol=1e-13)
    np.testing.assert_allclose(G, dq.geom_jacobian(base, q), atol=1e-13)


def test_batched_path_scores_match_oracle(dq):
    """Every agent's predicted path of a closed-loop rollout, scored on the device (one thread per path, nothing but
    the scores leaves the GPU), against the oracle's score of the same paths fetched to the host."""
    sc = scenarios.small_random(5, num_agents=70, num_obstacles=34, horizon=60)
    m = _planner()
    feed = loop.ObstacleFeed(sc)
    loop.plan_begin(m, sc)
    for _ in range(4):
        loop.control_tick(m, sc, feed)
        m.stop_prediction()
    lo, hi = m.panda_joint_limits()
    q0 = np.array([0.0, -0.4, 0.0, -2.0, 0.0, 1.6, 0.8])
    # an arm base from which the scene's start position is the end-effector position at q0
    base = IDENT.copy()
    base[5:] = 0.5 * (sc.start - dq.translation(dq.fkm(IDENT, q0)))
    paths = m.get_predicted_paths()
    steps = m.get_agent_summaries()["steps"]
    idx, got = m.score_paths(base, q0, k=0, damping=1e-3, tol_pos=2e-3)
    assert list(idx) == list(range(sc.num_agents))
    n_feasible = 0
    for a in range(sc.num_agents):
        want = dq.score_path(base, q0, paths[a, :steps[a]], lo, hi, damping=1e-3, tol_pos=2e-3)
        for k in ("max_pos_err", "min_joint_margin", "min_manipulability"):
            assert abs(got[k][a] - want[k]) <= 1e-9 * max(abs(want[k]), 1.0), (a, k, got[k][a], want[k])
        np.testing.assert_allclose(got["q_final"][a], want["q_final"], rtol=0, atol=1e-9)
        assert got["feasible"][a] == want["feasible"] and got["first_bad_point"][a] == want["first_bad_point"], a
        n_feasible += want["feasible"]
    assert 0 < n_feasible  # the case exercises feasible paths (and, with a tight tolerance, infeasible ones below)
    _, tight = m.score_paths(base, q0, k=0, damping=1e-3, tol_pos=1e-7)
    assert tight["feasible"].sum() < sc.num_agents
    # best-k: the same agents, in the same order, as the path export (pmaf_get_best_paths)
    best = m.evaluate_agents(feed.pos, feed.vel, feed.rad, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist, sc.k_workspace,
                             sc.ws_limits)
    kidx, kgot = m.score_paths(base, q0, k=5, damping=1e-3, tol_pos=2e-3)
    bidx = m.get_best_paths(5)[0]
    assert list(kidx) == list(bidx) and 0 <= best < sc.num_agents
    for r, a in enumerate(kidx):
        assert kgot["max_pos_err"][r] == got["max_pos_err"][a] and kgot["feasible"][r] == got["feasible"][a]
    m.close()
