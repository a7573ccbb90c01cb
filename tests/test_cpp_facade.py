"""The C++ drop-in (include/pmaf/cf_manager.hpp: the reference's CfManager class over the C ABI).
CPU: it compiles against an Eigen + Obstacle surface and links libpmaf.so. GPU: a C++ host program
driving it like the planner node reproduces the reference golden of the anchor task bit for bit."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from pmaf_b200 import planner, scenarios

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")


def _compile():
    planner.build()
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "dry_run")
    lib_dir = os.path.dirname(planner.LIB_PATH)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    subprocess.run([cxx, "-std=c++17", "-O2", "-ffp-contract=off", f"-I{ROOT}/include", f"-I{ROOT}/tests/cpp/stubs",
                    f"-I{ROOT}/oracle/shim", f"{ROOT}/tests/cpp/dry_run.cpp", "-o", exe, f"-L{lib_dir}", "-lpmaf",
                    f"-Wl,-rpath,{lib_dir}"], check=True)
    return exe


def test_facade_compiles_and_links():
    exe = _compile()
    assert os.path.exists(exe)
    # every CfManager member the planner node calls is declared
    hdr = open(os.path.join(ROOT, "include", "pmaf", "cf_manager.hpp")).read()
    for name in ("init", "setInitialPosition", "setRealEEAgentPosition", "stopPrediction", "evaluateAgents",
                 "getPredictedPaths", "moveRealEEAgent", "getNextPosition", "getNextVelocity", "resetEEAgents",
                 "startPrediction", "getDistFromGoal", "getPlannedTrajectory", "getGoalPosition", "getInitialPosition",
                 "getBestAgentType", "getPredictedPathLengths", "getPredictionTimes", "getAgentSuccess",
                 "getNumPredictionSteps", "joinPredictionThreads"):
        assert f" {name}(" in hdr, name


@pytest.mark.gpu
def test_cpp_host_program_matches_reference_golden():
    exe = _compile()
    sc = scenarios.anchor()
    vec_file = os.path.join(BUILD, "random_vecs.bin")
    np.ascontiguousarray(sc.random_vecs(), dtype=np.float64).tofile(vec_file)
    ticks = 40
    out = subprocess.run([exe, str(ticks), str(sc.max_prediction_steps), vec_file], check=True, capture_output=True,
                         text=True).stdout.strip().splitlines()
    want = np.load(os.path.join(ROOT, "tests", "golden", "anchor_A10_H1500.npz"))
    assert len(out) == ticks + 1
    for t in range(ticks):
        f = out[t].split()
        assert int(f[0]) == int(want["best"][t]), t
        got = np.array([float.fromhex(x) for x in f[1:7]])
        ref = np.concatenate([want["next_pos"][t], want["next_vel"][t]])
        assert np.array_equal(got, ref), (t, got, ref)
    tail = out[-1].split()
    assert int(tail[1]) == int(want["best_type"])
