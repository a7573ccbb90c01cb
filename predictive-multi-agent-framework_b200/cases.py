"""Named parity cases: each one drives an object with the CfManager call surface (the CUDA
product, the C oracle or the reference build) through a fixed script and returns the arrays a
parity test compares. tests/golden/make_golden.py runs them on the reference build to freeze
golden vectors; tests run them on the oracle (CPU) and on libpmaf.so (GPU).

The edge cases are the ones SURVEY.md §8c / Appendix A list for this path: O=1 and O=2, active
sentinel (repulsion), agent inside an obstacle (distance clamp 1e-5, +10000 cost), HAD rotation
vector NaN (d parallel to goal), goal inside approach_dist with early termination at 0.1, zero
relative velocity, obstacle on the shell boundary, moving obstacles re-fed every tick,
prediction_freq_multiple > 1, closed-loop position feedback, re-init on a live manager
(incumbent best agent survives init), per-agent gain jitter, ragged path lengths.
"""
from __future__ import annotations

import numpy as np

from . import loop, scenarios


def _finish(planner, rec, sc):
    known, rot = planner.get_obstacle_state()
    rec = dict(rec)
    rec["known"] = known
    rec["rot"] = rot
    rec["final_paths"] = planner.get_predicted_paths(sc.max_prediction_steps)
    rec["final_vel"] = planner.get_agent_velocities()
    rec["trajectory"] = planner.get_planned_trajectory()
    rec["best_type"] = np.array(planner.get_best_agent_type())
    rec["best_id"] = np.array(planner.get_best_agent_id())
    return rec


def closed_loop(sc, ticks):
    def run(planner):
        rec = loop.run_closed_loop(planner, sc, ticks)
        return _finish(planner, rec, sc)

    run.scenario = sc
    return run


def reinit(sc, ticks, new_goal):
    """Two `plan` goals on one manager: the incumbent best agent persists across init()."""

    def run(planner):
        rec1 = loop.run_closed_loop(planner, sc, ticks)
        sc2 = sc.with_(goal=np.array(new_goal, dtype=np.float64), start=planner.get_next_position(), seed=sc.seed + 1)
        rec2 = loop.run_closed_loop(planner, sc2, ticks)
        rec = {k: np.concatenate([rec1[k], rec2[k]]) for k in rec1}
        return _finish(planner, rec, sc2)

    run.scenario = sc
    return run


def feedback(sc, ticks, noise=1e-4):
    """open_loop: false — the measured position is pushed into the real agent every tick
    (panda_bimanual_control.cpp:333-335)."""

    def run(planner):
        rng = np.random.default_rng(sc.seed + 5)
        feed = loop.ObstacleFeed(sc)
        loop.plan_begin(planner, sc)
        best, pos, vel = [], [], []
        for _ in range(ticks):
            meas = planner.get_next_position() + noise * rng.uniform(-1, 1, 3)
            b, p, v = loop.control_tick(planner, sc, feed, measured_position=meas)
            best.append(b), pos.append(p), vel.append(v)
            feed.step()
        planner.stop_prediction()
        s = planner.get_agent_summaries()
        rec = dict(best=np.array(best), next_pos=np.array(pos), next_vel=np.array(vel), steps=s["steps"],
                   length=s["length"], min_obs_dist=s["min_obs_dist"], reached=s["reached"])
        return _finish(planner, rec, sc)

    run.scenario = sc
    return run


def zero_relative_velocity(sc):
    """Agents re-seeded with exactly the velocity of obstacle 0: |rel_vel| == 0 skips the circular
    force of an in-shell obstacle (cf_agent.cpp:97-98) while still latching its rotation vector."""

    def run(planner):
        feed = loop.ObstacleFeed(sc)
        loop.plan_begin(planner, sc)
        loop.control_tick(planner, sc, feed)
        planner.stop_prediction()
        v = np.array([0.15, 0.0, 0.0])
        feed.vel[0] = v
        p = feed.pos[0] - np.array([0.25, 0.0, 0.0])
        planner.reset_agents(p, v, feed.pos, feed.vel, feed.rad)
        planner.start_prediction()
        planner.stop_prediction()
        best = planner.evaluate_agents(feed.pos, feed.vel, feed.rad, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist,
                                       sc.k_workspace, sc.ws_limits)
        s = planner.get_agent_summaries()
        rec = dict(best=np.array([best]), steps=s["steps"], length=s["length"], min_obs_dist=s["min_obs_dist"],
                   reached=s["reached"])
        return _finish(planner, rec, sc)

    run.scenario = sc
    return run


def message_shaped_lists(sc, ticks):
    """Every call that takes an obstacle list gets the O-1 entries an `Obstacles.msg` carries — the obstacle
    node never publishes the sentinel (dynamic_obstacle_node.cpp:317,356) — instead of the planner's full
    list: moveRealEEAgent then treats the last PUBLISHED obstacle as the repulsion source (cf_agent.cpp:110-144,
    159-181 loop over the list they are given), resetEEAgents / setObstacles touch only the passed entries
    (cf_agent.cpp:63-70) and leave the agents' own sentinel and its known flag alone."""

    def run(planner):
        feed = loop.ObstacleFeed(sc)
        loop.plan_begin(planner, sc)
        rec = dict(best=[], next_pos=[], next_vel=[], steps=[], length=[], min_obs_dist=[], reached=[], goal_dist=[])
        for _ in range(ticks):
            op, ov, orad = feed.pos[:-1], feed.vel[:-1], feed.rad[:-1]
            planner.stop_prediction()
            best = planner.evaluate_agents(op, ov, orad, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist, sc.k_workspace,
                                           sc.ws_limits)
            planner.move_real_agent(op, ov, orad, sc.delta_t, 1, best)
            p, v = planner.get_next_position(), planner.get_next_velocity()
            planner.reset_agents(p, v, op, ov, orad)
            planner.start_prediction()
            planner.stop_prediction()
            s = planner.get_agent_summaries()
            rec["best"].append(best), rec["next_pos"].append(p), rec["next_vel"].append(v)
            for k in ("steps", "length", "min_obs_dist", "reached"):
                rec[k].append(s[k].copy())
            rec["goal_dist"].append(planner.get_dist_from_goal())
            feed.step()
        return _finish(planner, {k: np.array(v) for k, v in rec.items()}, sc)

    run.scenario = sc
    return run


def _with_obstacles(sc, pos, rad, vel=None, **kw):
    pos = np.array(list(pos) + [[100.0, 100.0, 100.0]], dtype=np.float64)
    rad = np.array(list(rad) + [0.1], dtype=np.float64)
    v = np.zeros_like(pos) if vel is None else np.array(list(vel) + [[0.0, 0.0, 0.0]], dtype=np.float64)
    return sc.with_(obs_pos=pos, obs_vel=v, obs_rad=rad, **kw)


def task_cases(golden_dir):
    """Closed-loop cases on the reference's own task-sequence files (config/tasks/*.yaml), replayed from the
    inputs stored in tests/golden/task_*.npz (written by make_golden.py where /root/reference exists)."""
    import glob
    import os

    out = {}
    for path in sorted(glob.glob(os.path.join(golden_dir, "task_*.npz"))):
        d = np.load(path)
        sc = scenarios.from_arrays(d)
        out[os.path.basename(path)[:-4]] = closed_loop(sc, int(d["in_ticks"]))
    return out


def all_cases(golden_dir=None):
    S = scenarios
    base = S.small_random(7, num_agents=8, num_obstacles=6, horizon=150)
    axis = base.with_(start=np.array([-0.7, 0.0, 0.65]), goal=np.array([0.5, 0.0, 0.65]))
    sentinel_near = S.small_random(11, num_agents=7, num_obstacles=5, horizon=100)
    sentinel_near.obs_pos[-1] = [-0.55, 0.12, 0.7]

    def near(seed):
        return S.small_random(seed, num_agents=14, num_obstacles=12, horizon=100).with_(
            start=np.array([-0.45, 0.0, 0.65]))

    c = {
        "anchor_A10_H1500": closed_loop(S.anchor(), 40),
        "anchor_A8_H50": closed_loop(S.anchor(8, 50), 20),
        "anchor_start2": closed_loop(S.anchor(10, 600, start=(-0.55, 0.05, 0.8), seed=3), 25),
        "rand0": closed_loop(S.small_random(0), 12),
        "rand1_jitter": closed_loop(S.small_random(1, gain_jitter=0.2), 12),
        "rand2_sparse": closed_loop(S.small_random(2, dense=False, num_obstacles=17, horizon=200), 15),
        "rand3_wide_shell": closed_loop(S.small_random(3, detect_shell_rad=0.6, velocity=0.35), 12),
        "rand4_long": closed_loop(S.small_random(4, num_agents=9, horizon=80), 120),
        "rand5_many_agents": closed_loop(S.small_random(5, num_agents=70, num_obstacles=34, horizon=60), 6),
        "rand6_many_obstacles": closed_loop(S.small_random(6, num_agents=6, num_obstacles=131, horizon=40, dense=False), 5),
        "moving0": closed_loop(S.small_random(100, moving=True), 15),
        "moving1": closed_loop(S.small_random(101, moving=True, num_obstacles=20, horizon=90), 15),
        "moving2_freq2": closed_loop(S.small_random(102, moving=True, prediction_freq_multiple=2), 10),
        "only_sentinel_O1": closed_loop(S.small_random(200, num_agents=6, num_obstacles=1, horizon=60), 5),
        "one_field_obstacle_O2": closed_loop(S.small_random(201, num_agents=6, num_obstacles=2, horizon=60), 8),
        "single_agent": closed_loop(S.small_random(202, num_agents=1, num_obstacles=5, horizon=60), 8),
        "sentinel_repels": closed_loop(sentinel_near, 15),
        "inside_obstacle": closed_loop(_with_obstacles(base, [[-0.7, 0.02, 0.65], [-0.3, 0.1, 0.7]], [0.08, 0.05]), 6),
        "had_nan_on_axis": closed_loop(_with_obstacles(axis, [[-0.3, 0.0, 0.65], [0.1, 0.2, 0.7]], [0.1, 0.05]), 40),
        "goal_inside_approach": closed_loop(base.with_(goal=np.array([-0.52, 0.05, 0.7])), 10),
        "shell_boundary": closed_loop(_with_obstacles(axis, [[-0.2, 0.0, 0.65], [-0.7, 0.5, 0.65]], [0.1, 0.1]), 10),
        "reinit_new_goal": reinit(S.anchor(10, 400), 12, [-0.2, 0.25, 0.9]),
        # start inside the obstacle field: the real agent latches obstacles and the best agent
        # switches between heuristic and RANDOM agents (seeds picked for that)
        "near326_switching": closed_loop(near(326), 80),
        "near312_switching": closed_loop(near(312), 60),
        "near301_clamped": closed_loop(near(301), 40),
        "near326_reinit_random_incumbent": reinit(near(326), 65, [-0.6, 0.2, 0.8]),
        "position_feedback": feedback(S.small_random(8, num_agents=8, horizon=100), 15),
        "zero_rel_velocity": zero_relative_velocity(S.small_random(9, num_agents=8, num_obstacles=4, horizon=50)),
        "message_lists_moving": message_shaped_lists(S.small_random(103, moving=True, num_agents=10, num_obstacles=11), 25),
        "message_lists_static": message_shaped_lists(near(326), 40),
    }
    if golden_dir is not None:
        c.update(task_cases(golden_dir))
    return c
