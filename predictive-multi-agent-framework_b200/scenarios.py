"""Workloads for the multi-agent predictive rollout: the BASELINE.json configs as concrete,
seeded inputs, plus a loader for the reference's task-sequence YAML keys.

All values are float64 numpy arrays laid out as the ROS wire types carry them
(`Position.msg` = float64[3], `Obstacles.msg` = Position[] pos, Position[] vel, float64[] radius;
/root/reference/src/bimanual_planning_ros/msg/Obstacles.msg:1-3). The LAST obstacle is the
self-collision sentinel (config/tasks/dual_arms_static1.yaml:66-69): it is the only source of
the repulsive force and is never published by the obstacle feed
(src/dynamic_obstacle_node.cpp:317,356).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace

import numpy as np


@dataclass
class Scenario:
    name: str
    num_agents: int
    goal: np.ndarray
    start: np.ndarray
    obs_pos: np.ndarray  # [O,3] incl. trailing sentinel
    obs_vel: np.ndarray  # [O,3]
    obs_rad: np.ndarray  # [O]
    k_attr: float = 4.0
    k_circ: float = 0.025
    k_repel: float = 0.08
    k_damp: float = 3.0
    k_manip: float = 0.0
    k_repel_body: float = 0.02
    k_goal_dist: float = 100.0
    k_path_len: float = 10.0
    k_safe_dist: float = 0.001
    k_workspace: float = 1.0
    ws_limits: np.ndarray = field(default_factory=lambda: np.array([1.0, -1.0, 0.3, -0.3, 1.1, 0.2]))
    max_prediction_steps: int = 1500
    approach_dist: float = 0.25
    detect_shell_rad: float = 0.35
    prediction_freq_multiple: int = 1
    frequency_ros: float = 100.0
    velocity: float = 0.2
    agent_mass: float = 1.0   # cf_manager.h:101 default, not forwarded by the node
    radius: float = 0.05      # cf_manager.h:102 default
    seed: int = 1
    feed_obstacles: bool = False  # advance obstacles 0..O-2 by vel/100 per tick (dynamic_obstacle_node.cpp:357)
    gain_jitter: float = 0.0      # optional +-fraction per-agent gain jitter (API takes per-agent vectors)

    @property
    def delta_t(self):
        return 1.0 / self.frequency_ros  # panda_bimanual_control.cpp:78-80

    @property
    def num_obstacles(self):
        return int(self.obs_rad.shape[0])

    def gains(self):
        """Per-agent gain vectors as the node passes them: uniform (panda_bimanual_control.cpp:463-471)."""
        A = self.num_agents
        out = {k: np.full(A, float(getattr(self, k))) for k in ("k_attr", "k_circ", "k_repel", "k_damp", "k_manip")}
        if self.gain_jitter > 0.0:
            rng = np.random.default_rng(self.seed + 77)
            for k in ("k_attr", "k_circ", "k_repel", "k_damp"):
                out[k] = out[k] * (1.0 + self.gain_jitter * rng.uniform(-1.0, 1.0, A))
        return out

    def random_vecs(self):
        """Seeded stand-in for RandomCfAgent's std::random_device draws (cf_agent.h:338-342):
        normalised U[-1,1]^3, agent-major / obstacle-minor, rows of heuristic agents unused."""
        rng = np.random.default_rng(self.seed)
        v = rng.uniform(-1.0, 1.0, (self.num_agents, self.num_obstacles, 3))
        n = np.sqrt((v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2])
        return v / n[..., None]

    def with_(self, **kw):
        return replace(self, **kw)


_SENTINEL = ([100.0, 100.0, 100.0], [0.0, 0.0, 0.0], 0.1)


def anchor(num_agents=10, max_prediction_steps=1500, start=(-0.6, 0.0, 0.65), seed=1):
    """C1: config/tasks/dual_arms_static1.yaml verbatim (A=10, O=10, H=1500); the start position is
    ours (the reference takes it from the simulator). BASELINE.json quotes 8 agents / horizon 50:
    pass num_agents=8, max_prediction_steps=50 for that variant."""
    pos = [[0.125, 0.0, 1.0], [0.125, 0.125, 1.0], [0.125, -0.125, 1.0],
           [0.125, 0.0, 0.7], [0.125, 0.125, 0.7], [0.125, -0.125, 0.7],
           [-0.35, 0.0, 0.6], [-0.35, 0.125, 0.6], [-0.35, -0.125, 0.6], _SENTINEL[0]]
    O = len(pos)
    return Scenario(name=f"dual_arms_static1_A{num_agents}_H{max_prediction_steps}", num_agents=num_agents,
                    goal=np.array([0.5, 0.0, 0.7]), start=np.array(start, dtype=np.float64),
                    obs_pos=np.array(pos), obs_vel=np.zeros((O, 3)), obs_rad=np.full(O, 0.1),
                    max_prediction_steps=max_prediction_steps, seed=seed)


def synthetic(name, num_agents, num_obstacles, horizon, seed, velocity=0.2, moving=False, **kw):
    """C2-C5 generator: obstacles 0..O-2 uniform in [-0.6,0.8]x[-0.5,0.5]x[0.2,1.1],
    radius U[0.03,0.08], sentinel last; start (-0.7,0,0.65); goal far enough that no agent
    terminates early, so every one of the A*(H-1) integration steps executes."""
    rng = np.random.default_rng(seed)
    O = num_obstacles
    pos = np.empty((O, 3))
    pos[:, 0] = rng.uniform(-0.6, 0.8, O)
    pos[:, 1] = rng.uniform(-0.5, 0.5, O)
    pos[:, 2] = rng.uniform(0.2, 1.1, O)
    rad = rng.uniform(0.03, 0.08, O)
    vel = rng.uniform(-0.1, 0.1, (O, 3)) if moving else np.zeros((O, 3))
    pos[-1], vel[-1], rad[-1] = _SENTINEL
    dt = 1.0 / 100.0
    goal = np.array([-0.7 + velocity * dt * horizon + 0.5, 0.0, 0.7])
    return Scenario(name=name, num_agents=num_agents, goal=goal, start=np.array([-0.7, 0.0, 0.65]),
                    obs_pos=pos, obs_vel=vel, obs_rad=rad, max_prediction_steps=horizon, velocity=velocity,
                    seed=seed, feed_obstacles=moving, **kw)


def c2():
    """256 agents, 64 obstacles, horizon 200 — the configuration BASELINE.json's metric is quoted on."""
    return synthetic("c2_256x64x200", 256, 64, 200, seed=1002)


def c3():
    """4096 agents, 256 obstacles, horizon 500 (roofline capture)."""
    return synthetic("c3_4096x256x500", 4096, 256, 500, seed=1003)


def c4(num_agents=65536):
    """65536 agents, 1024 obstacles, horizon 200, agents sharded over 8 GPUs."""
    return synthetic(f"c4_{num_agents}x1024x200", num_agents, 1024, 200, seed=1004)


def c5():
    """Kobo dual-arm: 1024 agents, 50 moving obstacles, horizon 300, gains from
    config/tasks/sim_kobo_dyn_spheres1.yaml:4-20, obstacles re-fed every tick."""
    return synthetic("c5_kobo_1024x50x300", 1024, 50, 300, seed=1005, velocity=0.3, moving=True,
                     k_attr=8.0, k_circ=0.10, k_repel=0.08, k_damp=5.0,
                     ws_limits=np.array([0.8, 0.35, 0.6, -0.6, 0.8, 0.2]))


def small_random(seed, num_agents=12, num_obstacles=9, horizon=120, moving=False, dense=True, **kw):
    """Small randomised parity case: obstacles packed around the start->goal line so that every
    agent type meets in-shell obstacles within a few steps."""
    rng = np.random.default_rng(seed)
    O = num_obstacles
    pos = np.empty((O, 3))
    span = (0.9, 0.5, 0.5) if dense else (1.4, 1.0, 0.9)
    pos[:, 0] = rng.uniform(-0.45, -0.45 + span[0], O)
    pos[:, 1] = rng.uniform(-span[1] / 2, span[1] / 2, O)
    pos[:, 2] = rng.uniform(0.65 - span[2] / 2, 0.65 + span[2] / 2, O)
    rad = rng.uniform(0.03, 0.1, O)
    vel = rng.uniform(-0.15, 0.15, (O, 3)) if moving else np.zeros((O, 3))
    pos[-1], vel[-1], rad[-1] = _SENTINEL
    return Scenario(name=f"rand{seed}_{num_agents}x{O}x{horizon}", num_agents=num_agents,
                    goal=np.array([0.5, 0.0, 0.7]), start=np.array([-0.7, 0.02, 0.65]),
                    obs_pos=pos, obs_vel=vel, obs_rad=rad, max_prediction_steps=horizon, seed=seed,
                    feed_obstacles=moving, **kw)


PLANNER_KEYS = ("num_agents_ee", "k_attr", "k_circ", "k_repel", "k_damp", "k_manip", "k_repel_body", "k_goal_dist",
                "k_path_len", "k_safe_dist", "k_workspace", "desired_ws_limits", "max_prediction_steps",
                "approach_dist", "detect_shell_rad", "prediction_freq_multiple", "frequency_ros", "velocity")


def from_task_yaml(path, start, goal_index=None, seed=1):
    """Load a reference task-sequence file (config/tasks/*.yaml) — the keys the planner consumes
    (panda_bimanual_control.cpp:129-161; list in SURVEY.md §8b). The goal is the first `plan`
    entry of `goals:` (or goals[goal_index])."""
    import yaml

    with open(path) as f:
        cfg = yaml.safe_load(f)["bimanual_planning"]
    obs = cfg["obstacles"]
    goals = cfg["goals"]
    plan = goals[goal_index] if goal_index is not None else next(g for g in goals if g.get("type") == "plan")
    feed = any(any(float(x) != 0.0 for x in o["vel"]) for o in obs[:-1])
    return Scenario(
        name=path.rsplit("/", 1)[-1].rsplit(".", 1)[0], num_agents=int(cfg["num_agents_ee"]),
        goal=np.array(plan["pos"], dtype=np.float64), start=np.array(start, dtype=np.float64),
        obs_pos=np.array([o["pos"] for o in obs], dtype=np.float64),
        obs_vel=np.array([o["vel"] for o in obs], dtype=np.float64),
        obs_rad=np.array([o["radius"] for o in obs], dtype=np.float64),
        k_attr=float(cfg["k_attr"]), k_circ=float(cfg["k_circ"]), k_repel=float(cfg["k_repel"]),
        k_damp=float(cfg["k_damp"]), k_manip=float(cfg["k_manip"]), k_repel_body=float(cfg["k_repel_body"]),
        k_goal_dist=float(cfg["k_goal_dist"]), k_path_len=float(cfg["k_path_len"]),
        k_safe_dist=float(cfg["k_safe_dist"]), k_workspace=float(cfg["k_workspace"]),
        ws_limits=np.array(cfg["desired_ws_limits"], dtype=np.float64),
        max_prediction_steps=int(cfg["max_prediction_steps"]), approach_dist=float(cfg["approach_dist"]),
        detect_shell_rad=float(cfg["detect_shell_rad"]),
        prediction_freq_multiple=int(cfg["prediction_freq_multiple"]), frequency_ros=float(cfg["frequency_ros"]),
        velocity=float(cfg["velocity"]), seed=seed, feed_obstacles=feed)


_SCALARS = ("name", "num_agents", "k_attr", "k_circ", "k_repel", "k_damp", "k_manip", "k_repel_body", "k_goal_dist",
            "k_path_len", "k_safe_dist", "k_workspace", "max_prediction_steps", "approach_dist", "detect_shell_rad",
            "prediction_freq_multiple", "frequency_ros", "velocity", "agent_mass", "radius", "seed", "feed_obstacles",
            "gain_jitter")
_ARRAYS = ("goal", "start", "obs_pos", "obs_vel", "obs_rad", "ws_limits")


def to_arrays(sc, prefix="in_"):
    """Flatten a Scenario into numpy arrays (stored inside golden files so that a case built from a
    reference task YAML can be replayed where /root/reference does not exist)."""
    d = {prefix + k: np.array(getattr(sc, k)) for k in _SCALARS}
    d.update({prefix + k: np.array(getattr(sc, k), dtype=np.float64) for k in _ARRAYS})
    return d


def from_arrays(d, prefix="in_"):
    kw = {}
    for k in _SCALARS:
        v = d[prefix + k]
        v = v.item() if hasattr(v, "item") else v
        kw[k] = str(v) if k == "name" else v
    for k in ("num_agents", "max_prediction_steps", "prediction_freq_multiple", "seed"):
        kw[k] = int(kw[k])
    kw["feed_obstacles"] = bool(kw["feed_obstacles"])
    for k in _ARRAYS:
        kw[k] = np.array(d[prefix + k], dtype=np.float64)
    return Scenario(**kw)
