"""In-process stand-in for the reference's `dry_run.launch` relay
(/root/reference/src/bimanual_planning_ros/launch/dry_run.launch:9,41: the planner's `goals`
output is relayed back as its `position` input) driving any object with the CfManager call
surface through the control-tick order of `planCallback`
(src/panda_bimanual_control.cpp:329-369) and `taskCallback` (:494-521)."""
from __future__ import annotations

import numpy as np

from .scenarios import Scenario


class ObstacleFeed:
    """`dynamic_obstacle_node` main loop (src/dynamic_obstacle_node.cpp:352-369): obstacles
    0..O-2 advance by vel/frequency per feed tick, velocities stay as configured; the message
    omits the sentinel, so `obstacleCallback` (src/panda_bimanual_control.cpp:302-309) leaves the
    planner's last obstacle untouched."""

    def __init__(self, sc: Scenario, frequency=100.0):
        self.pos = sc.obs_pos.copy()
        self.vel = sc.obs_vel.copy()
        self.rad = sc.obs_rad.copy()
        self.frequency = float(frequency)
        self.active = bool(sc.feed_obstacles)

    def step(self):
        if self.active:
            self.pos[:-1] += self.vel[:-1] / self.frequency


def plan_begin(planner, sc: Scenario, random_vecs=True):
    """`plan` goal activation (taskCallback, :494-521): init -> setInitialPosition."""
    g = sc.gains()
    planner.init(sc.goal, sc.delta_t, sc.obs_pos, sc.obs_vel, sc.obs_rad, g["k_attr"], g["k_circ"], g["k_repel"],
                 g["k_damp"], g["k_manip"], [sc.k_repel_body], sc.velocity, sc.approach_dist, sc.detect_shell_rad,
                 sc.max_prediction_steps, sc.prediction_freq_multiple, sc.agent_mass, sc.radius)
    if random_vecs:
        planner.set_random_vecs(sc.random_vecs())
    planner.set_initial_position(sc.start)


def control_tick(planner, sc: Scenario, feed: ObstacleFeed, measured_position=None):
    """One `planCallback` (:329-369). Returns the best agent index and the published goal."""
    if measured_position is not None:  # open_loop: false (:333-335)
        planner.set_real_position(measured_position)
    planner.stop_prediction()
    best = planner.evaluate_agents(feed.pos, feed.vel, feed.rad, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist,
                                   sc.k_workspace, sc.ws_limits)
    planner.move_real_agent(feed.pos, feed.vel, feed.rad, sc.delta_t, 1, best)
    nxt_p, nxt_v = planner.get_next_position(), planner.get_next_velocity()
    planner.reset_agents(nxt_p, nxt_v, feed.pos, feed.vel, feed.rad)
    planner.start_prediction()
    return best, nxt_p, nxt_v


def run_closed_loop(planner, sc: Scenario, ticks, record_paths=False, begin=True):
    """Closed dry-run loop; returns per-tick records (what a parity test compares)."""
    feed = ObstacleFeed(sc)
    if begin:
        plan_begin(planner, sc)
    rec = dict(best=[], next_pos=[], next_vel=[], steps=[], length=[], min_obs_dist=[], reached=[], paths=[],
               goal_dist=[])
    for _ in range(ticks):
        best, p, v = control_tick(planner, sc, feed)
        planner.stop_prediction()  # wait for the rollout this tick started so that it can be read
        s = planner.get_agent_summaries()
        rec["best"].append(best)
        rec["next_pos"].append(p)
        rec["next_vel"].append(v)
        rec["steps"].append(s["steps"].copy())
        rec["length"].append(s["length"].copy())
        rec["min_obs_dist"].append(s["min_obs_dist"].copy())
        rec["reached"].append(s["reached"].copy())
        rec["goal_dist"].append(planner.get_dist_from_goal())
        if record_paths:
            rec["paths"].append(planner.get_predicted_paths())
        feed.step()
    out = {k: np.array(v) for k, v in rec.items() if len(v)}
    return out
