"""Host-side mirror of the reference's planner API over libpmaf.so (include/pmaf.h).

`CfManager` exposes the methods the planner node calls on
ghostplanner::cfplanner::CfManager (/root/reference/src/bimanual_planning_ros/include/
bimanual_planning_ros/cf_manager.h:36-136; call sites in src/panda_bimanual_control.cpp:329-369,
:463-471, :494-521) with snake_case names, numpy arrays for Eigen vectors / Obstacle lists, and
the reference's error behaviour mapped to exceptions (`PmafError`, carrying the C status).

There is no CPU implementation behind this class: if libpmaf.so is missing, or no sm_100 GPU is
visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PMAF_LIB: developer override to load an instrumented variant of the library (tools/*.py)
LIB_PATH = os.path.join(_HERE, os.environ.get("PMAF_LIB", "libpmaf.so"))

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class PmafError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"pmaf status {status}: {message}")
        self.status = status


class PathScore(C.Structure):
    _fields_ = [("max_pos_err", C.c_double), ("min_joint_margin", C.c_double), ("min_manipulability", C.c_double),
                ("feasible", C.c_int), ("first_bad_point", C.c_int), ("q_final", C.c_double * 7)]


class Counters(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("collectives", C.c_uint64), ("rollouts", C.c_uint64), ("agent_steps", C.c_uint64), ("agent_steps_total", C.c_uint64),
                ("last_rollout_ms", C.c_double), ("rollout_ms_total", C.c_double), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("lanes_per_agent", C.c_int), ("block_threads", C.c_int),
                ("grid_blocks", C.c_int), ("smem_bytes", C.c_int), ("occupancy_build", C.c_int), ("reserved_", C.c_int),
                ("general_steps_total", C.c_uint64)]


# every symbol include/pmaf.h declares (tests check that the library exports all of them)
API_SYMBOLS = [
    "pmaf_last_error", "pmaf_version", "pmaf_create", "pmaf_destroy", "pmaf_set_shard", "pmaf_nccl_unique_id", "pmaf_nccl_init", "pmaf_set_nccl_comm", "pmaf_p2p_export", "pmaf_p2p_import",
    "pmaf_init", "pmaf_seed_random_vecs", "pmaf_set_random_vecs", "pmaf_get_random_vecs",
    "pmaf_set_initial_position", "pmaf_set_real_position", "pmaf_start_prediction", "pmaf_stop_prediction",
    "pmaf_evaluate_agents", "pmaf_move_real_agent", "pmaf_reset_agents", "pmaf_feed_obstacles", "pmaf_tick", "pmaf_get_num_agents",
    "pmaf_get_next_position", "pmaf_get_next_velocity", "pmaf_get_ee_force", "pmaf_get_goal_position",
    "pmaf_get_initial_position", "pmaf_get_dist_from_goal", "pmaf_get_best_agent_type", "pmaf_get_best_agent_id",
    "pmaf_get_num_prediction_steps", "pmaf_get_real_num_prediction_steps", "pmaf_get_agent_summaries",
    "pmaf_get_predicted_paths", "pmaf_get_predicted_path", "pmaf_get_agent_velocities",
    "pmaf_get_planned_trajectory", "pmaf_get_obstacle_state", "pmaf_get_costs", "pmaf_get_counters", "pmaf_get_fast_stats", "pmaf_dry_run",
    "pmaf_set_tuning", "pmaf_set_upload_dedup", "pmaf_set_rollout_timing", "pmaf_timer_start", "pmaf_timer_stop",
    "pmaf_flush_l2", "pmaf_measure_fp64_peak", "pmaf_selftest_math", "pmaf_get_section_cycles", "pmaf_get_best_paths",
    "pmaf_dq_kinematics", "pmaf_panda_joint_limits", "pmaf_score_paths",
]


def build(verbose=False):
    """Compile libpmaf.so in-tree for sm_100a (csrc/Makefile)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building libpmaf.so failed")
    return LIB_PATH


_lib = None


def load_library():
    """dlopen libpmaf.so and declare its prototypes. Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.pmaf_last_error.restype = C.c_char_p
    lib.pmaf_version.restype = C.c_char_p
    lib.pmaf_create.argtypes = [C.POINTER(H), C.c_int]
    lib.pmaf_destroy.argtypes = [H]
    lib.pmaf_set_shard.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.pmaf_set_nccl_comm.argtypes = [H, C.c_void_p]
    lib.pmaf_nccl_unique_id.argtypes = [C.c_char_p]
    lib.pmaf_nccl_init.argtypes = [H, C.c_char_p, C.c_int, C.c_int]
    lib.pmaf_p2p_export.argtypes = [H, C.c_char_p]
    lib.pmaf_p2p_import.argtypes = [H, C.c_char_p, C.c_int, C.c_int]
    lib.pmaf_get_section_cycles.argtypes = [H, C.POINTER(C.c_longlong)]
    lib.pmaf_init.argtypes = [H, _dp, C.c_double, C.c_int, _dp, _dp, _dp, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_int,
                              _dp, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.c_double, C.c_double]
    lib.pmaf_seed_random_vecs.argtypes = [H, C.c_uint64]
    lib.pmaf_set_random_vecs.argtypes = [H, _dp, C.c_int, C.c_int]
    lib.pmaf_get_random_vecs.argtypes = [H, _dp, C.c_int, C.c_int]
    lib.pmaf_set_initial_position.argtypes = [H, _dp]
    lib.pmaf_set_real_position.argtypes = [H, _dp]
    lib.pmaf_start_prediction.argtypes = [H]
    lib.pmaf_stop_prediction.argtypes = [H]
    lib.pmaf_evaluate_agents.argtypes = [H, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double,
                                         _dp, _ip]
    lib.pmaf_move_real_agent.argtypes = [H, C.c_int, _dp, _dp, _dp, C.c_double, C.c_int, C.c_int]
    lib.pmaf_reset_agents.argtypes = [H, _dp, _dp, C.c_int, _dp, _dp, _dp]
    lib.pmaf_feed_obstacles.argtypes = [H, C.c_int, C.c_double]
    lib.pmaf_tick.argtypes = [H, _dp, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double,
                              C.c_double, _dp, _ip, _dp, _dp]
    lib.pmaf_get_num_agents.argtypes = [H, _ip]
    for n in ("pmaf_get_next_position", "pmaf_get_next_velocity", "pmaf_get_ee_force", "pmaf_get_goal_position",
              "pmaf_get_initial_position", "pmaf_get_dist_from_goal"):
        getattr(lib, n).argtypes = [H, _dp]
    lib.pmaf_get_best_agent_type.argtypes = [H, _ip]
    lib.pmaf_get_best_agent_id.argtypes = [H, _ip]
    lib.pmaf_get_num_prediction_steps.argtypes = [H, C.c_int, _ip]
    lib.pmaf_get_real_num_prediction_steps.argtypes = [H, _ip]
    lib.pmaf_get_agent_summaries.argtypes = [H, _ip, _dp, _dp, _ip, _dp, _ip]
    lib.pmaf_get_predicted_paths.argtypes = [H, _dp, C.c_int]
    lib.pmaf_get_predicted_path.argtypes = [H, C.c_int, _dp, C.c_int, _ip]
    lib.pmaf_get_agent_velocities.argtypes = [H, _dp]
    lib.pmaf_get_planned_trajectory.argtypes = [H, _dp, C.c_int, _ip]
    lib.pmaf_get_obstacle_state.argtypes = [H, C.c_int, _ip, _dp]
    lib.pmaf_get_costs.argtypes = [H, _dp]
    lib.pmaf_get_best_paths.argtypes = [H, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp]
    lib.pmaf_get_counters.argtypes = [H, C.POINTER(Counters)]
    lib.pmaf_get_fast_stats.argtypes = [H, C.POINTER(C.c_uint64)]
    lib.pmaf_dry_run.argtypes = [H, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double,
                                 C.c_double, C.c_double, C.c_double, _dp, C.c_int, _dp, _ip, _dp, _dp]
    lib.pmaf_dq_kinematics.argtypes = [H, _dp, _dp, _dp, _dp, _dp]
    lib.pmaf_panda_joint_limits.argtypes = [_dp, _dp]
    lib.pmaf_panda_joint_limits.restype = None
    lib.pmaf_score_paths.argtypes = [H, C.c_int, _dp, _dp, _dp, _dp, C.c_double, C.c_double, _ip, C.POINTER(PathScore)]
    lib.pmaf_set_tuning.argtypes = [H, C.c_int, C.c_int, C.c_int]
    lib.pmaf_set_upload_dedup.argtypes = [H, C.c_int]
    lib.pmaf_set_rollout_timing.argtypes = [H, C.c_int]
    lib.pmaf_timer_start.argtypes = [H]
    lib.pmaf_timer_stop.argtypes = [H, _dp]
    lib.pmaf_flush_l2.argtypes = [H]
    lib.pmaf_measure_fp64_peak.argtypes = [H, _dp]
    lib.pmaf_selftest_math.argtypes = [H, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    _lib = lib
    return lib


def _f64(x, shape=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    return a.reshape(shape) if shape is not None else a


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class CfManager:
    """ghostplanner::cfplanner::CfManager over the C ABI. One instance = one planner on one GPU."""

    def __init__(self, device=0, lanes_per_agent=0, block_threads=0, occupancy=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        self._check(self.lib.pmaf_create(C.byref(self.h), int(device)))
        if lanes_per_agent or block_threads or occupancy:
            self._check(self.lib.pmaf_set_tuning(self.h, int(lanes_per_agent), int(block_threads), int(occupancy)))
        self.A = self.O = self.H = 0
        self.n_global = 0
        self.first_agent = 0

    def _check(self, rc):
        if rc != 0:
            raise PmafError(rc, self.lib.pmaf_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.pmaf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration ----------------------------------------------------------------------
    def set_shard(self, n_global, first_agent, rank, world):
        self._check(self.lib.pmaf_set_shard(self.h, n_global, first_agent, rank, world))
        self.first_agent = first_agent

    def seed_random_vecs(self, seed):
        self._check(self.lib.pmaf_seed_random_vecs(self.h, int(seed)))

    def set_tuning(self, lanes_per_agent=0, block_threads=0, occupancy=0):
        self._check(self.lib.pmaf_set_tuning(self.h, int(lanes_per_agent), int(block_threads), int(occupancy)))

    def init(self, goal, delta_t, obs_pos, obs_vel, obs_rad, k_attr, k_circ, k_repel, k_damp, k_manip,
             k_repel_force=(), velocity_max=0.5, approach_dist=0.25, detect_shell_rad=0.8,
             max_prediction_steps=1500, prediction_freq_multiple=1, agent_mass=1.0, radius=0.05):
        op, ov, orad = _f64(obs_pos, (-1, 3)), _f64(obs_vel, (-1, 3)), _f64(obs_rad, (-1,))
        ka, kc, kr, kd, km = (_f64(x, (-1,)) for x in (k_attr, k_circ, k_repel, k_damp, k_manip))
        kf = _f64(k_repel_force, (-1,))
        # the reference asserts k_attr/k_circ/k_repel/k_manip (cf_manager.cpp:50-51) and reads k_damp[i] for every
        # agent as well (:73-104): a shorter k_damp is an out-of-bounds read there, an error here
        if not (len(ka) == len(kc) == len(kr) == len(km)) or len(kd) < len(ka):
            raise PmafError(-1, "gain vectors differ in length")
        self._check(self.lib.pmaf_init(self.h, _d(_f64(goal, (3,))), float(delta_t), len(orad), _d(op), _d(ov),
                                       _d(orad), len(ka), _d(ka), _d(kc), _d(kr), _d(kd), _d(km), len(kf), _d(kf),
                                       float(velocity_max), float(approach_dist), float(detect_shell_rad),
                                       int(max_prediction_steps), int(prediction_freq_multiple), float(agent_mass),
                                       float(radius)))
        n = C.c_int()
        self._check(self.lib.pmaf_get_num_agents(self.h, C.byref(n)))
        self.A, self.O, self.H, self.n_global = n.value, len(orad), int(max_prediction_steps), max(len(ka), 1)

    def set_random_vecs(self, vecs):
        v = _f64(vecs)
        self._check(self.lib.pmaf_set_random_vecs(self.h, _d(v), v.shape[0], v.shape[1]))

    def get_random_vecs(self):
        v = np.zeros((self.A, self.O, 3))
        self._check(self.lib.pmaf_get_random_vecs(self.h, _d(v), self.A, self.O))
        return v

    # ---- the per-tick calls (planCallback order) ---------------------------------------------------
    def set_initial_position(self, p):
        self._check(self.lib.pmaf_set_initial_position(self.h, _d(_f64(p, (3,)))))

    def set_real_position(self, p):
        self._check(self.lib.pmaf_set_real_position(self.h, _d(_f64(p, (3,)))))

    def start_prediction(self):
        self._check(self.lib.pmaf_start_prediction(self.h))

    def stop_prediction(self):
        self._check(self.lib.pmaf_stop_prediction(self.h))

    def evaluate_agents(self, obs_pos, obs_vel, obs_rad, k_goal_dist, k_path_len, k_safe_dist, k_workspace,
                        ws_limits):
        op, ov, orad = _f64(obs_pos, (-1, 3)), _f64(obs_vel, (-1, 3)), _f64(obs_rad, (-1,))
        best = C.c_int()
        self._check(self.lib.pmaf_evaluate_agents(self.h, len(orad), _d(op), _d(ov), _d(orad), float(k_goal_dist),
                                                  float(k_path_len), float(k_safe_dist), float(k_workspace),
                                                  _d(_f64(ws_limits, (6,))), C.byref(best)))
        return best.value

    def move_real_agent(self, obs_pos, obs_vel, obs_rad, delta_t, steps, agent_id):
        op, ov, orad = _f64(obs_pos, (-1, 3)), _f64(obs_vel, (-1, 3)), _f64(obs_rad, (-1,))
        self._check(self.lib.pmaf_move_real_agent(self.h, len(orad), _d(op), _d(ov), _d(orad), float(delta_t),
                                                  int(steps), int(agent_id)))

    def reset_agents(self, pos, vel, obs_pos, obs_vel, obs_rad):
        op, ov, orad = _f64(obs_pos, (-1, 3)), _f64(obs_vel, (-1, 3)), _f64(obs_rad, (-1,))
        self._check(self.lib.pmaf_reset_agents(self.h, _d(_f64(pos, (3,))), _d(_f64(vel, (3,))), len(orad), _d(op),
                                               _d(ov), _d(orad)))

    def tick(self, obs_pos, obs_vel, obs_rad, delta_t, k_goal_dist, k_path_len, k_safe_dist, k_workspace, ws_limits,
             measured_position=None):
        """One whole planCallback as a device-resident chain (pmaf_tick). obs_pos = obs_vel = obs_rad = None: the
        device-resident list as the last upload / feed_obstacles left it."""
        best = C.c_int()
        p, v = np.zeros(3), np.zeros(3)
        meas = _d(_f64(measured_position, (3,))) if measured_position is not None else None
        if obs_pos is None:
            self._check(self.lib.pmaf_tick(self.h, meas, self.O, None, None, None, float(delta_t), float(k_goal_dist),
                                           float(k_path_len), float(k_safe_dist), float(k_workspace),
                                           _d(_f64(ws_limits, (6,))), C.byref(best), _d(p), _d(v)))
            return best.value, p, v
        op, ov, orad = _f64(obs_pos, (-1, 3)), _f64(obs_vel, (-1, 3)), _f64(obs_rad, (-1,))
        self._check(self.lib.pmaf_tick(self.h, meas, len(orad), _d(op), _d(ov), _d(orad), float(delta_t),
                                       float(k_goal_dist), float(k_path_len), float(k_safe_dist), float(k_workspace),
                                       _d(_f64(ws_limits, (6,))), C.byref(best), _d(p), _d(v)))
        return best.value, p, v

    def feed_obstacles(self, n_feed, frequency=100.0):
        """The obstacle node's integration step on the device-resident live list (pmaf_feed_obstacles)."""
        self._check(self.lib.pmaf_feed_obstacles(self.h, int(n_feed), float(frequency)))

    def dry_run(self, ticks, obs_pos, obs_vel, obs_rad, n_feed, delta_t, k_goal_dist, k_path_len, k_safe_dist,
                k_workspace, ws_limits, feed_frequency=100.0, wait_rollout=False, flush_l2=False, profile=None,
                tick_times=None, device_feed=False):
        """`ticks` planCallbacks in the library's C++ host loop (pmaf_dry_run) on HOST obstacle arrays; obs_pos is
        advanced in place by the obstacle feed. Returns (seconds inside the ticks, best[ticks], next_pos, next_vel)."""
        assert obs_pos.dtype == np.float64 and obs_pos.flags["C_CONTIGUOUS"]
        ov, orad = _f64(obs_vel, (-1, 3)), _f64(obs_rad, (-1,))
        best = np.zeros(max(ticks, 1), dtype=np.int32)
        npos, nvel = np.zeros((max(ticks, 1), 3)), np.zeros((max(ticks, 1), 3))
        sec = (C.c_double * (7 + max(ticks, 0)))()
        flags = (1 if wait_rollout else 0) | (2 if flush_l2 else 0) | (4 if profile is not None else 0) | \
            (8 if tick_times is not None else 0) | (16 if device_feed else 0)
        self._check(self.lib.pmaf_dry_run(self.h, int(ticks), len(orad), _d(obs_pos), _d(ov), _d(orad), int(n_feed),
                                          float(feed_frequency), float(delta_t), float(k_goal_dist), float(k_path_len),
                                          float(k_safe_dist), float(k_workspace), _d(_f64(ws_limits, (6,))), flags,
                                          sec, _i(best), _d(npos), _d(nvel)))
        if profile is not None:  # per-call wall time: stop, evaluate, move_real, get + reset, start, final wait
            profile[:] = [sec[k] for k in range(1, 7)]
        if tick_times is not None:  # wall time of every tick
            tick_times[:] = [sec[7 + t] for t in range(ticks)]
        return sec[0], best[:ticks], npos[:ticks], nvel[:ticks]

    # ---- getters ----------------------------------------------------------------------------------
    def _vec3(self, fn):
        v = np.zeros(3)
        self._check(fn(self.h, _d(v)))
        return v

    def get_next_position(self):
        return self._vec3(self.lib.pmaf_get_next_position)

    def get_next_velocity(self):
        return self._vec3(self.lib.pmaf_get_next_velocity)

    def get_ee_force(self):
        return self._vec3(self.lib.pmaf_get_ee_force)

    def get_goal_position(self):
        return self._vec3(self.lib.pmaf_get_goal_position)

    def get_initial_position(self):
        return self._vec3(self.lib.pmaf_get_initial_position)

    def get_dist_from_goal(self):
        d = C.c_double()
        self._check(self.lib.pmaf_get_dist_from_goal(self.h, C.byref(d)))
        return d.value

    def get_best_agent_type(self):
        t = C.c_int()
        self._check(self.lib.pmaf_get_best_agent_type(self.h, C.byref(t)))
        return t.value

    def get_best_agent_id(self):
        t = C.c_int()
        self._check(self.lib.pmaf_get_best_agent_id(self.h, C.byref(t)))
        return t.value

    def get_num_prediction_steps(self, agent):
        t = C.c_int()
        self._check(self.lib.pmaf_get_num_prediction_steps(self.h, int(agent), C.byref(t)))
        return t.value

    def get_agent_summaries(self):
        A = self.A
        steps, reached, types = (np.zeros(A, dtype=np.int32) for _ in range(3))
        length, mind, t = (np.zeros(A) for _ in range(3))
        self._check(self.lib.pmaf_get_agent_summaries(self.h, _i(steps), _d(length), _d(mind), _i(reached), _d(t),
                                                      _i(types)))
        return dict(steps=steps, length=length, min_obs_dist=mind, reached=reached, pred_time_ns=t, agent_type=types)

    def get_predicted_paths(self, stride=None):
        stride = int(stride or self.H)
        out = np.full((self.A, stride, 3), np.nan)
        self._check(self.lib.pmaf_get_predicted_paths(self.h, _d(out), stride))
        return out

    def get_predicted_path(self, agent):
        out = np.zeros((self.H, 3))
        n = C.c_int()
        self._check(self.lib.pmaf_get_predicted_path(self.h, int(agent), _d(out), self.H, C.byref(n)))
        return out[: n.value]

    def get_agent_velocities(self):
        out = np.zeros((self.A, 3))
        self._check(self.lib.pmaf_get_agent_velocities(self.h, _d(out)))
        return out

    def get_planned_trajectory(self):
        n = C.c_int()
        self._check(self.lib.pmaf_get_planned_trajectory(self.h, None, 0, C.byref(n)))
        out = np.zeros((max(n.value, 1), 3))
        self._check(self.lib.pmaf_get_planned_trajectory(self.h, _d(out), n.value, C.byref(n)))
        return out[: n.value]

    def get_obstacle_state(self):
        known = np.zeros((self.A + 1, self.O), dtype=np.int32)
        rot = np.zeros((self.A + 1, self.O, 3))
        self._check(self.lib.pmaf_get_obstacle_state(self.h, self.O, _i(known), _d(rot)))
        return known, rot

    def get_costs(self):
        c = np.zeros(self.A)
        self._check(self.lib.pmaf_get_costs(self.h, _d(c)))
        return c

    def counters(self):
        c = Counters()
        self._check(self.lib.pmaf_get_counters(self.h, C.byref(c)))
        return {f[0]: getattr(c, f[0]) for f in Counters._fields_}

    def fast_stats(self):
        """Reasons that kept steps off the straight-line step (PMAF_FAST_STATS builds; zeros otherwise)."""
        self.counters()
        out = (C.c_uint64 * 12)()
        self._check(self.lib.pmaf_get_fast_stats(self.h, out))
        names = ["unusable_or_candidates", "start_threshold", "range_distance_chain", "first_detection_general",
                 "range_force_chain", "force_threshold", "sentinel_in_reach", "acceleration_clamp", "range_integrator"]
        return {n: int(out[i]) for i, n in enumerate(names)}

    def set_rollout_timing(self, on):
        """CUDA events around every rollout kernel (counters' rollout times); off = production setting."""
        self._check(self.lib.pmaf_set_rollout_timing(self.h, int(bool(on))))

    def set_upload_dedup(self, dedup):
        self._check(self.lib.pmaf_set_upload_dedup(self.h, 1 if dedup else 0))

    def timer_start(self):
        self._check(self.lib.pmaf_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self._check(self.lib.pmaf_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        self._check(self.lib.pmaf_flush_l2(self.h))

    def measure_fp64_peak(self):
        t = C.c_double()
        self._check(self.lib.pmaf_measure_fp64_peak(self.h, C.byref(t)))
        return t.value

    def selftest_math(self, samples, seed=1):
        out = (C.c_uint64 * 5)()
        self._check(self.lib.pmaf_selftest_math(self.h, int(samples), int(seed), out))
        return dict(zip(("sqrt_mismatch", "div_mismatch", "div3_mismatch", "flagged", "compared"), list(out)))

    # ---- downstream kinematics ------------------------------------------------
    def dq_kinematics(self, base_dq, q):
        """Pose (8), pose Jacobian (8 x 7) and geometric Jacobian (6 x 7) of the Panda at q, on the device."""
        pose, J, G = np.zeros(8), np.zeros((8, 7)), np.zeros((6, 7))
        self._check(self.lib.pmaf_dq_kinematics(self.h, _d(_f64(base_dq, (8,))), _d(_f64(q, (7,))), _d(pose), _d(J), _d(G)))
        return pose, J, G

    def panda_joint_limits(self):
        lo, hi = np.zeros(7), np.zeros(7)
        self.lib.pmaf_panda_joint_limits(_d(lo), _d(hi))
        return lo, hi

    def score_paths(self, base_dq, q_start, k=0, q_lo=None, q_hi=None, damping=1e-3, tol_pos=1e-3):
        """Feasibility scores of the predicted paths of the last rollout (k = 0: every local agent; k >= 1: the k
        cheapest agents of the last evaluate). Returns (agent_index[n], dict of arrays)."""
        lo, hi = self.panda_joint_limits()
        lo = _f64(q_lo, (7,)) if q_lo is not None else lo
        hi = _f64(q_hi, (7,)) if q_hi is not None else hi
        n = k if k > 0 else self.A
        idx = np.zeros(n, dtype=np.int32)
        out = (PathScore * n)()
        self._check(self.lib.pmaf_score_paths(self.h, int(k), _d(_f64(base_dq, (8,))), _d(_f64(q_start, (7,))), _d(lo), _d(hi),
                                              float(damping), float(tol_pos), _i(idx), out))
        rec = dict(max_pos_err=np.array([o.max_pos_err for o in out]), min_joint_margin=np.array([o.min_joint_margin for o in out]),
                   min_manipulability=np.array([o.min_manipulability for o in out]),
                   feasible=np.array([o.feasible for o in out]), first_bad_point=np.array([o.first_bad_point for o in out]),
                   q_final=np.array([list(o.q_final) for o in out]))
        return idx, rec

    def get_best_paths(self, k, stride=1, max_points=None):
        """The k cheapest agents of the last evaluate and their (decimated) paths."""
        max_points = int(max_points or self.H)
        idx, n = np.zeros(k, dtype=np.int32), np.zeros(k, dtype=np.int32)
        paths = np.full((k, max_points, 3), np.nan)
        self._check(self.lib.pmaf_get_best_paths(self.h, int(k), int(stride), max_points, _i(idx), _i(n), _d(paths)))
        return idx, n, paths
