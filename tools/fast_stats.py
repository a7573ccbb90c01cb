"""Which steps miss the straight-line step, and why? Runs a workload on the statistics build
(libpmaf_stats.so: csrc compiled with -DPMAF_FAST_STATS) and prints the reason counts.

    make -C predictive-multi-agent-framework_b200/csrc OUT=../libpmaf_stats.so EXTRA=-DPMAF_FAST_STATS
    python tools/fast_stats.py [c2|c5|anchor] [ticks]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pmaf_b200  # noqa: E402,F401
from pmaf_b200 import loop, planner, scenarios  # noqa: E402

if __name__ == "__main__":
    planner.LIB_PATH = os.path.join(os.path.dirname(planner.LIB_PATH), "libpmaf_stats.so")
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    sc = scenarios.anchor(10, 1500) if name == "anchor" else getattr(scenarios, name)()
    m = planner.CfManager(0)
    feed = loop.ObstacleFeed(sc)
    loop.plan_begin(m, sc)
    for _ in range(ticks):
        loop.control_tick(m, sc, feed)
        feed.step()
    m.stop_prediction()
    c = m.counters()
    print(f"{sc.name}: {c['general_steps_total']} of {c['agent_steps_total']} steps took the general step; "
          f"last rollout {c['last_rollout_ms']:.3f} ms")
    for k, v in m.fast_stats().items():
        print(f"  {k:28s} {v}")
    import ctypes as C
    import numpy as np
    m.lib.pmaf_get_section_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
    out = np.zeros((64, 12), dtype=np.int64)
    m.lib.pmaf_get_section_cycles(m.h, out.ctypes.data_as(C.POINTER(C.c_longlong)))
    print("  per agent (first 16 of the last rollout): fast cycles/step, general cycles/step, general steps, candidates/step")
    for a in range(min(16, sc.num_agents)):
        f, g_, ng, nc, n, lc, ln = out[a, :7]
        nf = max(n - ng, 1)
        print(f"   agent {a:2d}: fast {f / nf:7.0f}  general {g_ / max(ng, 1):7.0f}  general steps {ng:4d} of {n:4d}  candidates/step {nc / max(n, 1):5.1f}  cold latches {ln:3d} x {lc / max(ln, 1):6.0f} cycles")
    m.close()
