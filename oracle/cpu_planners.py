"""TEST INFRASTRUCTURE — ctypes front-ends for the two CPU checkers.

* ``RefPlanner``    -> oracle/_ref/libcfref.so : the reference's own CfManager/CfAgent code
                       (compiled unmodified from /root/reference by oracle/Makefile).
* ``OraclePlanner`` -> oracle/libcforacle.so   : cf_oracle.c, the plain-C restatement.

Both expose the call sequence the planner node makes on ``CfManager``
(/root/reference/src/bimanual_planning_ros/src/panda_bimanual_control.cpp:329-369) with numpy
arrays, and the same method names as the product's host mirror so that a parity test can drive
all three with one script.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module. The product (libpmaf.so and the package around it) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libcfref.so")
ORACLE_LIB = os.path.join(_HERE, "libcforacle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f64(x, shape=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if shape is not None:
        a = a.reshape(shape)
    return a


def build(target="all"):
    """(Re)build the checkers; `ref` is skipped by the Makefile when /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", _HERE, target], check=True)


def have_ref():
    return os.path.exists(REF_LIB)


def have_oracle():
    return os.path.exists(ORACLE_LIB)


class _CpuPlanner:
    """Shared ctypes plumbing; subclasses set PREFIX and LIB."""

    PREFIX = ""
    LIB = ""
    _lib_cache = {}

    def __init__(self, threads=0, pooled=True):
        if self.LIB not in self._lib_cache:
            lib = C.CDLL(self.LIB)
            self._declare(lib)
            self._lib_cache[self.LIB] = lib
        self.lib = self._lib_cache[self.LIB]
        self.h = C.c_void_p(self._fn("create")())
        self.threads = int(threads)
        self.pooled = bool(pooled)
        self.A = 0
        self.O = 0
        self.H = 0
        self.last_rollout_seconds = 0.0

    def _fn(self, name):
        return getattr(self.lib, f"{self.PREFIX}_{name}")

    def _declare(self, lib):
        p = self.PREFIX
        g = lambda n: getattr(lib, f"{p}_{n}")
        g("create").restype = C.c_void_p
        g("destroy").argtypes = [C.c_void_p]
        g("init").argtypes = [C.c_void_p, _dp, C.c_double, C.c_int, _dp, _dp, _dp, C.c_int, _dp, _dp,
                              _dp, _dp, _dp, C.c_int, _dp, C.c_double, C.c_double, C.c_double,
                              C.c_ulong, C.c_ulong, C.c_double, C.c_double, C.c_int]
        g("num_agents").argtypes = [C.c_void_p]
        g("set_random_vecs").argtypes = [C.c_void_p, _dp, C.c_int]
        g("get_random_vecs").argtypes = [C.c_void_p, _dp, C.c_int]
        g("set_initial_position").argtypes = [C.c_void_p, _dp]
        g("set_real_position").argtypes = [C.c_void_p, _dp]
        g("start_prediction").argtypes = [C.c_void_p]
        g("stop_prediction").argtypes = [C.c_void_p]
        g("rollout_threads").argtypes = [C.c_void_p]
        g("rollout_threads").restype = C.c_double
        g("rollout_pooled").argtypes = [C.c_void_p, C.c_int]
        g("rollout_pooled").restype = C.c_double
        g("evaluate_agents").argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double,
                                         C.c_double, C.c_double, _dp]
        g("evaluate_agents").restype = C.c_int
        g("move_real_agent").argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, C.c_double, C.c_int, C.c_int]
        g("reset_agents").argtypes = [C.c_void_p, _dp, _dp, C.c_int, _dp, _dp, _dp]
        for n in ("get_next_position", "get_next_velocity", "get_ee_force"):
            g(n).argtypes = [C.c_void_p, _dp]
        g("get_dist_from_goal").argtypes = [C.c_void_p]
        g("get_dist_from_goal").restype = C.c_double
        g("get_best_agent_type").argtypes = [C.c_void_p]
        g("get_best_agent_id").argtypes = [C.c_void_p]
        g("get_num_prediction_steps").argtypes = [C.c_void_p, C.c_int]
        g("get_real_num_steps").argtypes = [C.c_void_p]
        g("get_agent_summaries").argtypes = [C.c_void_p, _ip, _dp, _dp, _ip, _dp, _ip]
        g("get_predicted_paths").argtypes = [C.c_void_p, _dp, C.c_int]
        g("get_agent_velocities").argtypes = [C.c_void_p, _dp]
        g("get_planned_trajectory").argtypes = [C.c_void_p, _dp, C.c_int]
        g("get_planned_trajectory").restype = C.c_int
        g("get_obstacle_state").argtypes = [C.c_void_p, C.c_int, _ip, _dp]
        g("host_threads").restype = C.c_int

    def close(self):
        if self.h:
            self._fn("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- CfManager surface -------------------------------------------------
    def init(self, goal, delta_t, obs_pos, obs_vel, obs_rad, k_attr, k_circ, k_repel, k_damp, k_manip,
             k_repel_force=(), velocity_max=0.5, approach_dist=0.25, detect_shell_rad=0.8,
             max_prediction_steps=1500, prediction_freq_multiple=1, agent_mass=1.0, radius=0.05):
        op, ov, orad = _f64(obs_pos, (-1, 3)), _f64(obs_vel, (-1, 3)), _f64(obs_rad, (-1,))
        ka, kc, kr, kd, km = (_f64(x, (-1,)) for x in (k_attr, k_circ, k_repel, k_damp, k_manip))
        kf = _f64(k_repel_force, (-1,))
        self.A, self.O, self.H = len(ka), len(orad), int(max_prediction_steps)
        self._fn("init")(self.h, _d(_f64(goal)), float(delta_t), self.O, _d(op), _d(ov), _d(orad), self.A,
                         _d(ka), _d(kc), _d(kr), _d(kd), _d(km), len(kf), _d(kf), float(velocity_max),
                         float(approach_dist), float(detect_shell_rad), self.H,
                         int(prediction_freq_multiple), float(agent_mass), float(radius),
                         0 if self.pooled else 1)

    def set_random_vecs(self, vecs):
        v = _f64(vecs, (self.A, self.O, 3))
        self._fn("set_random_vecs")(self.h, _d(v), self.O)

    def get_random_vecs(self):
        v = np.zeros((self.A, self.O, 3))
        self._fn("get_random_vecs")(self.h, _d(v), self.O)
        return v

    def set_initial_position(self, p):
        self._fn("set_initial_position")(self.h, _d(_f64(p)))

    def set_real_position(self, p):
        self._fn("set_real_position")(self.h, _d(_f64(p)))

    def start_prediction(self):
        """Run every agent's rollout to termination (see ref_harness.cpp on anytime semantics)."""
        if self.pooled:
            self.last_rollout_seconds = self._fn("rollout_pooled")(self.h, self.threads)
        else:
            self.last_rollout_seconds = self._fn("rollout_threads")(self.h)

    def stop_prediction(self):
        if not self.pooled:
            self._fn("stop_prediction")(self.h)

    def evaluate_agents(self, obs_pos, obs_vel, obs_rad, k_goal_dist, k_path_len, k_safe_dist, k_workspace,
                        ws_limits):
        op, ov, orad = _f64(obs_pos, (-1, 3)), _f64(obs_vel, (-1, 3)), _f64(obs_rad, (-1,))
        return self._fn("evaluate_agents")(self.h, len(orad), _d(op), _d(ov), _d(orad), float(k_goal_dist),
                                           float(k_path_len), float(k_safe_dist), float(k_workspace),
                                           _d(_f64(ws_limits, (6,))))

    def move_real_agent(self, obs_pos, obs_vel, obs_rad, delta_t, steps, agent_id):
        op, ov, orad = _f64(obs_pos, (-1, 3)), _f64(obs_vel, (-1, 3)), _f64(obs_rad, (-1,))
        self._fn("move_real_agent")(self.h, len(orad), _d(op), _d(ov), _d(orad), float(delta_t), int(steps),
                                    int(agent_id))

    def reset_agents(self, pos, vel, obs_pos, obs_vel, obs_rad):
        op, ov, orad = _f64(obs_pos, (-1, 3)), _f64(obs_vel, (-1, 3)), _f64(obs_rad, (-1,))
        self._fn("reset_agents")(self.h, _d(_f64(pos)), _d(_f64(vel)), len(orad), _d(op), _d(ov), _d(orad))

    # ---- getters -------------------------------------------------------------
    def _vec3(self, name):
        v = np.zeros(3)
        self._fn(name)(self.h, _d(v))
        return v

    def get_next_position(self):
        return self._vec3("get_next_position")

    def get_next_velocity(self):
        return self._vec3("get_next_velocity")

    def get_dist_from_goal(self):
        return self._fn("get_dist_from_goal")(self.h)

    def get_best_agent_type(self):
        return self._fn("get_best_agent_type")(self.h)

    def get_best_agent_id(self):
        return self._fn("get_best_agent_id")(self.h)

    def get_agent_summaries(self):
        A = self.A
        steps, reached, types = (np.zeros(A, dtype=np.int32) for _ in range(3))
        length, mind, t = (np.zeros(A) for _ in range(3))
        self._fn("get_agent_summaries")(self.h, _i(steps), _d(length), _d(mind), _i(reached), _d(t), _i(types))
        return dict(steps=steps, length=length, min_obs_dist=mind, reached=reached, pred_time_ns=t,
                    agent_type=types)

    def get_predicted_paths(self, stride=None):
        stride = int(stride or self.H)
        out = np.full((self.A, stride, 3), np.nan)
        self._fn("get_predicted_paths")(self.h, _d(out), stride)
        return out

    def get_agent_velocities(self):
        out = np.zeros((self.A, 3))
        self._fn("get_agent_velocities")(self.h, _d(out))
        return out

    def get_planned_trajectory(self):
        n = self._fn("get_real_num_steps")(self.h)
        out = np.zeros((max(n, 1), 3))
        self._fn("get_planned_trajectory")(self.h, _d(out), n)
        return out[:n]

    def get_obstacle_state(self):
        known = np.zeros((self.A + 1, self.O), dtype=np.int32)
        rot = np.zeros((self.A + 1, self.O, 3))
        self._fn("get_obstacle_state")(self.h, self.O, _i(known), _d(rot))
        return known, rot

    def host_threads(self):
        """Threads the pooled driver runs on: the explicit count, else OpenMP's default (which a launcher
        may have pinned to 1 through OMP_NUM_THREADS — pass `threads` to override it)."""
        return self.threads if self.threads > 0 else self._fn("host_threads")()


class RefPlanner(_CpuPlanner):
    PREFIX = "cfref"
    LIB = REF_LIB


class OraclePlanner(_CpuPlanner):
    PREFIX = "cforacle"
    LIB = ORACLE_LIB


# ---- downstream dual-quaternion kinematics (dq_oracle.c, PARITY UNPINNED) -----------------
DQ_LIB = os.path.join(_HERE, "libdqoracle.so")


class DqPathScore(C.Structure):
    _fields_ = [("max_pos_err", C.c_double), ("min_joint_margin", C.c_double), ("min_manipulability", C.c_double),
                ("feasible", C.c_int), ("first_bad_point", C.c_int), ("q_final", C.c_double * 7)]


class DqOracle:
    """ctypes front end of oracle/dq_oracle.c."""

    def __init__(self):
        if not os.path.exists(DQ_LIB):
            build("oracle")
        self.lib = C.CDLL(DQ_LIB)
        self.lib.dqo_fkm.argtypes = [_dp, _dp, _dp]
        self.lib.dqo_pose_jacobian.argtypes = [_dp, _dp, _dp]
        self.lib.dqo_translation.argtypes = [_dp, _dp]
        self.lib.dqo_geom_jacobian.argtypes = [_dp, _dp, _dp]
        self.lib.dqo_score_path.argtypes = [_dp, _dp, _dp, C.c_int, _dp, _dp, C.c_double, C.c_double, C.POINTER(DqPathScore)]

    def fkm(self, base, q):
        out = np.zeros(8)
        self.lib.dqo_fkm(_d(_f64(base, (8,))), _d(_f64(q, (7,))), _d(out))
        return out

    def pose_jacobian(self, base, q):
        out = np.zeros((8, 7))
        self.lib.dqo_pose_jacobian(_d(_f64(base, (8,))), _d(_f64(q, (7,))), _d(out))
        return out

    def translation(self, x):
        out = np.zeros(3)
        self.lib.dqo_translation(_d(_f64(x, (8,))), _d(out))
        return out

    def geom_jacobian(self, base, q):
        out = np.zeros((6, 7))
        self.lib.dqo_geom_jacobian(_d(_f64(base, (8,))), _d(_f64(q, (7,))), _d(out))
        return out

    def score_path(self, base, q_start, path, q_lo, q_hi, damping=1e-3, tol_pos=1e-3):
        path = _f64(path, (-1, 3))
        out = DqPathScore()
        self.lib.dqo_score_path(_d(_f64(base, (8,))), _d(_f64(q_start, (7,))), _d(path), len(path), _d(_f64(q_lo, (7,))),
                                _d(_f64(q_hi, (7,))), float(damping), float(tol_pos), C.byref(out))
        return dict(max_pos_err=out.max_pos_err, min_joint_margin=out.min_joint_margin,
                    min_manipulability=out.min_manipulability, feasible=out.feasible,
                    first_bad_point=out.first_bad_point, q_final=np.array(list(out.q_final)))
