"""Where do the cycles of one agent-step of the GENERAL step go? Runs a workload on the instrumented build
(libpmaf_timers.so: make -C .../csrc OUT=../libpmaf_timers.so EXTRA=-DPMAF_SECTION_TIMERS) and prints cycles
per step and section. The section markers sit in agent_step / field_pass; the straight-line steps of the
latency build (pmaf_fast.cuh) are profiled with tools/fast_stats.py and ncu instead, so only the steps that
fall back to the general step are accounted here.

    python tools/section_timers.py [c2|c3|c5]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pmaf_b200  # noqa: E402,F401
from pmaf_b200 import loop, planner, scenarios  # noqa: E402

SECTIONS = ["0 prologue (norms, desired velocity)", "1 broad phase", "2 unit vectors", "3 candidate evaluation",
            "4 commit + ordered force sum", "5 reductions", "6 repel/attract/integrate", "7 path store + counters",
            "8 gate", "9 segment + workspace cost"]

if __name__ == "__main__":
    planner.LIB_PATH = os.path.join(os.path.dirname(planner.LIB_PATH), "libpmaf_timers.so")
    sc = getattr(scenarios, sys.argv[1] if len(sys.argv) > 1 else "c2")()
    m = planner.CfManager(0)
    m.lib.pmaf_get_section_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
    feed = loop.ObstacleFeed(sc)
    loop.plan_begin(m, sc)
    for _ in range(6):
        loop.control_tick(m, sc, feed)
    m.stop_prediction()
    out = np.zeros((64, 12), dtype=np.int64)
    m.lib.pmaf_get_section_cycles(m.h, out.ctypes.data_as(C.POINTER(C.c_longlong)))
    steps = m.get_agent_summaries()["steps"][:64] - 1
    per_step = out[:, :10] / np.maximum(steps, 1)[:, None]
    print(f"{sc.name}: cycles per agent-step, mean over the first 64 agents (last rollout); kernel "
          f"{m.counters()['last_rollout_ms']:.3f} ms")
    for name, v in zip(SECTIONS, per_step.mean(0)):
        print(f"  {name:42s} {v:8.0f}")
    print(f"  {'total':42s} {per_step.sum(1).mean():8.0f}")
