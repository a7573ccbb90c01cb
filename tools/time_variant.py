"""Developer tool: time a workload's rollout kernel on a library variant (e.g. libpmaf_exp.so).
    python tools/time_variant.py libpmaf_exp.so c2 [ticks]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pmaf_b200  # noqa: E402,F401
from pmaf_b200 import loop, planner, scenarios  # noqa: E402

if __name__ == "__main__":
    planner.LIB_PATH = os.path.join(os.path.dirname(planner.LIB_PATH), sys.argv[1])
    sc = getattr(scenarios, sys.argv[2])()
    ticks = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    m = planner.CfManager(0)
    feed = loop.ObstacleFeed(sc)
    loop.plan_begin(m, sc)
    ms = []
    for t in range(ticks):
        m.flush_l2()
        loop.control_tick(m, sc, feed)
        feed.step()
        m.stop_prediction()
        ms.append(m.counters()["last_rollout_ms"])
    c = m.counters()
    ms = sorted(ms[5:])
    print(f"{sys.argv[1]} {sc.name}: rollout median {ms[len(ms)//2]:.4f} ms  min {ms[0]:.4f}  general {c['general_steps_total']}/{c['agent_steps_total']}")
    m.close()
