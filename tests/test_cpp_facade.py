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


REF_PKG = "/root/reference/src/bimanual_planning_ros"


def test_reference_planner_node_compiles_against_the_facade():
    """f1's claim, exercised: the reference's planner node — src/panda_bimanual_control.cpp, UNMODIFIED, read where
    it lies — passes the compiler's full semantic analysis (`g++ -fsyntax-only`: every CfManager call the node makes
    is resolved against include/pmaf/cf_manager.hpp) with the one-line replacement of cf_manager.h that
    INTEGRATION.md §1 describes. ROS, dynamic_reconfigure, actionlib, the generated messages and Eigen are absent
    from the image: tests/cpp/ros_stubs/ declares their surface (declarations only, nothing is linked or run).
    Needs the reference checkout, so it runs in the dev container only."""
    src = os.path.join(REF_PKG, "src", "panda_bimanual_control.cpp")
    if not os.path.exists(src):
        pytest.skip("reference sources not present")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    cmd = [cxx, "-std=c++17", "-fsyntax-only", "-w", "-H", f"-I{ROOT}/tests/cpp/ros_stubs", f"-I{ROOT}/include",
           f"-I{ROOT}/tests/cpp/stubs", f"-I{REF_PKG}/include", src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    included = r.stderr
    assert "include/pmaf/cf_manager.hpp" in included and "include/pmaf.h" in included  # the facade, not ...
    assert f"{REF_PKG}/include/bimanual_planning_ros/cf_manager.h" not in included    # ... the reference's class
    assert f"{REF_PKG}/include/bimanual_planning_ros/cf_agent.h" not in included
    assert f"{REF_PKG}/include/bimanual_planning_ros/panda_bimanual_control.h" in included  # the node's own header is the reference's
    # and the check bites: without a member the node calls, the same command fails
    broken = os.path.join(BUILD, "broken_facade")
    os.makedirs(os.path.join(broken, "pmaf"), exist_ok=True)
    hdr = open(os.path.join(ROOT, "include", "pmaf", "cf_manager.hpp")).read()
    assert "  void resetEEAgents(" in hdr
    open(os.path.join(broken, "pmaf", "cf_manager.hpp"), "w").write(hdr.replace("  void resetEEAgents(", "  void resetEEAgentsRenamed("))
    shutil.copy(os.path.join(ROOT, "include", "pmaf.h"), os.path.join(broken, "pmaf.h"))
    bad = subprocess.run([c if c != f"-I{ROOT}/include" else f"-I{broken}" for c in cmd], capture_output=True, text=True)
    assert bad.returncode != 0 and "resetEEAgents" in bad.stderr


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
