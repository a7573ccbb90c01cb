"""CPU check of the CUDA source's operation order: the product's step logic (agent_step /
field_pass in csrc/*.cuh) compiled for the HOST with a single-lane group policy must reproduce
the oracle's rollouts bit for bit, tick after tick, from the oracle's own per-agent state.
(The warp-parallel execution of the same source is what the -m gpu tests check.)"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from pmaf_b200 import cases, loop, scenarios

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
_dp = C.POINTER(C.c_double)


def _d(a):
    return a.ctypes.data_as(_dp)


@pytest.fixture(scope="module")
def hoststep():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "libhoststep.so")
    src = os.path.join(HERE, "host_step_check.cu")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
                    "-gencode", "arch=compute_100a,code=sm_100a", "-o", so, src], check=True)
    lib = C.CDLL(so)
    lib.hoststep_rollout.restype = C.c_int
    return lib


ORDER = [6, 1, 2, 3, 4]  # agent type by index (cf_manager.cpp:70-104)


def _margin(sc):
    s = max(np.abs(sc.goal).max(), np.abs(sc.start).max(), np.abs(sc.obs_pos).max())
    s += (sc.velocity + np.abs(sc.obs_vel).max() * 1.7320508) * sc.delta_t * sc.prediction_freq_multiple * sc.max_prediction_steps
    return np.float32(1e-3 + 4e-6 * s)


def _check_scenario(lib, oracle, sc, ticks):
    feed = loop.ObstacleFeed(sc)
    loop.plan_begin(oracle, sc)
    gains = sc.gains()
    rnd = sc.random_vecs()
    A, O, H = sc.num_agents, sc.num_obstacles, sc.max_prediction_steps
    dt = sc.delta_t * sc.prediction_freq_multiple
    for t in range(ticks):
        # planCallback up to the reset, on the oracle
        oracle.stop_prediction()
        best = oracle.evaluate_agents(feed.pos, feed.vel, feed.rad, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist,
                                      sc.k_workspace, sc.ws_limits)
        oracle.move_real_agent(feed.pos, feed.vel, feed.rad, sc.delta_t, 1, best)
        p0, v0 = oracle.get_next_position(), oracle.get_next_velocity()
        oracle.reset_agents(p0, v0, feed.pos, feed.vel, feed.rad)
        known0, rot0 = oracle.get_obstacle_state()
        oracle.start_prediction()
        want_paths = oracle.get_predicted_paths()
        summ = oracle.get_agent_summaries()
        known1, rot1 = oracle.get_obstacle_state()
        want_vel = oracle.get_agent_velocities()
        vn = np.sqrt((v0[0] * v0[0] + v0[1] * v0[1]) + v0[2] * v0[2])
        v0c = v0 * (sc.velocity / vn) if vn > sc.velocity else v0
        for a in range(A):
            known = known0[a].astype(np.uint8).copy()
            rot = rot0[a].copy()
            path = np.full((H, 3), np.nan)
            path[0] = p0
            n_path = C.c_int(1)
            v_out = np.zeros(3)
            mo, pl = C.c_double(), C.c_double()
            lib.hoststep_rollout(
                O, _d(feed.pos), _d(feed.vel), _d(sc.obs_rad), _d(sc.goal), C.c_double(sc.detect_shell_rad),
                C.c_double(sc.agent_mass), C.c_double(sc.radius), C.c_double(sc.velocity), C.c_double(sc.approach_dist),
                C.c_double(dt), H, ORDER[a] if a < 5 else 5, C.c_double(gains["k_attr"][a]),
                C.c_double(gains["k_circ"][a]), C.c_double(gains["k_repel"][a]), C.c_double(gains["k_damp"][a]),
                _d(np.ascontiguousarray(sc.start)), _d(p0), _d(np.ascontiguousarray(v0c)),
                C.c_double(sc.detect_shell_rad), known.ctypes.data_as(C.POINTER(C.c_ubyte)), _d(rot),
                _d(np.ascontiguousarray(rnd[a])), C.c_float(_margin(sc)), _d(path), C.byref(n_path), _d(v_out),
                C.byref(mo), C.byref(pl))
            ctx = f"{sc.name} tick {t} agent {a}"
            assert n_path.value == summ["steps"][a], ctx
            n = n_path.value
            assert np.array_equal(path[:n], want_paths[a, :n], equal_nan=True), ctx
            assert np.array_equal(v_out, want_vel[a], equal_nan=True), ctx
            assert mo.value == summ["min_obs_dist"][a], ctx
            assert pl.value == summ["length"][a] or (np.isnan(pl.value) and np.isnan(summ["length"][a])), ctx
            assert np.array_equal(known, known1[a]), ctx
            assert np.array_equal(rot, rot1[a], equal_nan=True), ctx
        feed.step()


CASES = cases.all_cases()
NAMES = ["anchor_A8_H50", "rand0", "rand1_jitter", "rand3_wide_shell", "rand6_many_obstacles", "moving0",
         "moving2_freq2", "only_sentinel_O1", "one_field_obstacle_O2", "sentinel_repels", "inside_obstacle",
         "had_nan_on_axis", "goal_inside_approach", "shell_boundary", "near326_switching", "near301_clamped"]


@pytest.mark.parametrize("name", NAMES)
def test_cuda_step_source_matches_oracle_on_host(name, hoststep, oracle_built):
    sc = CASES[name].scenario
    ticks = {"near326_switching": 70, "had_nan_on_axis": 40, "near301_clamped": 40}.get(name, 8)
    _check_scenario(hoststep, oracle_built.OraclePlanner(), sc, ticks)


def test_cuda_step_source_on_anchor(hoststep, oracle_built):
    _check_scenario(hoststep, oracle_built.OraclePlanner(), scenarios.anchor(), 6)
