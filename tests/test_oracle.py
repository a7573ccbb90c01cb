"""CPU tests of the oracle itself: the plain-C restatement (oracle/cf_oracle.c)
must reproduce, bit for bit, (a) the golden vectors frozen from the reference build and (b) the
reference build itself, when it is available, on fresh randomised cases."""
import glob
import hashlib
import os

import numpy as np
import pytest

from pmaf_b200 import cases, loop, scenarios
from parity import assert_bit_identical

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = cases.all_cases(GOLDEN)


def _digest(sc):
    h = hashlib.sha256()
    for a in (sc.goal, sc.start, sc.obs_pos, sc.obs_vel, sc.obs_rad, sc.random_vecs(), *sc.gains().values()):
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


def test_every_case_has_a_golden_file():
    have = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))}
    assert have == set(CASES), have ^ set(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_golden(name, oracle_built):
    want = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    assert str(want["sha_inputs"]) == _digest(CASES[name].scenario), "case inputs drifted from the golden file"
    p = oracle_built.OraclePlanner()
    got = CASES[name](p)
    p.close()
    assert_bit_identical(got, want, ctx=f"{name}: ")


@pytest.mark.parametrize("seed", range(400, 406))
def test_oracle_matches_reference_build_on_fresh_cases(seed, oracle_built):
    if not oracle_built.have_ref():
        pytest.skip("reference build (oracle/_ref) not present")
    sc = scenarios.small_random(seed, num_agents=11, num_obstacles=10, horizon=90, moving=bool(seed % 2),
                                gain_jitter=0.1 * (seed % 3)).with_(start=np.array([-0.5, 0.01 * (seed % 5), 0.66]))
    ref, orc = oracle_built.RefPlanner(), oracle_built.OraclePlanner()
    a = loop.run_closed_loop(ref, sc, 30, record_paths=True)
    b = loop.run_closed_loop(orc, sc, 30, record_paths=True)
    assert_bit_identical(b, a, ctx=f"seed {seed}: ")
    ka, ra = ref.get_obstacle_state()
    kb, rb = orc.get_obstacle_state()
    assert_bit_identical(dict(known=kb, rot=rb), dict(known=ka, rot=ra))
    ref.close(), orc.close()


def test_reference_thread_driver_equals_pooled_driver(oracle_built):
    """The reference's own thread-per-agent driver (cf_manager.cpp:118-123), run to termination,
    gives the same result as the pooled driver over its per-step methods (ref_harness.cpp)."""
    if not oracle_built.have_ref():
        pytest.skip("reference build (oracle/_ref) not present")
    sc = scenarios.anchor(8, 300)
    a = loop.run_closed_loop(oracle_built.RefPlanner(pooled=False), sc, 6, record_paths=True)
    b = loop.run_closed_loop(oracle_built.RefPlanner(pooled=True), sc, 6, record_paths=True)
    assert_bit_identical(a, b)


def test_first_tick_picks_agent_zero_and_anchor_switches(oracle_built):
    """Quirk 7: every path has one point on the first tick, all costs are
    equal and index 0 (HAD) wins; on the anchor task GOAL_OBSTACLE (index 3) takes over."""
    p = oracle_built.OraclePlanner()
    rec = loop.run_closed_loop(p, scenarios.anchor(), 5)
    assert rec["best"][0] == 0 and rec["best"][1] == 3
    np.testing.assert_allclose(rec["next_pos"][0], [-0.599871531, 0.0, 0.650001362], atol=5e-10)
