"""How long does each agent's rollout take? The rollout kernel's time is the slowest agent's (one warp per
agent in the latency build), so stragglers matter as much as the median. Prints the distribution of the
per-agent rollout times the kernel records (prediction_time_, cf_agent.cpp:329-331) on a workload.

    python tools/per_agent_time.py [c2|c5] [ticks]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pmaf_b200  # noqa: E402,F401
from pmaf_b200 import loop, planner, scenarios  # noqa: E402

if __name__ == "__main__":
    sc = getattr(scenarios, sys.argv[1] if len(sys.argv) > 1 else "c2")()
    ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    m = planner.CfManager(0)
    feed = loop.ObstacleFeed(sc)
    loop.plan_begin(m, sc)
    for _ in range(ticks):
        m.flush_l2()
        loop.control_tick(m, sc, feed)
        feed.step()
        m.stop_prediction()
    t = m.get_agent_summaries()["pred_time_ns"] / 1e3
    print(f"rollout ms {m.counters()['last_rollout_ms']:.4f}")
    print(f"per-agent us: min {t.min():.1f} med {np.median(t):.1f} p90 {np.percentile(t, 90):.1f} max {t.max():.1f}")
    order = np.argsort(-t)
    print("slowest:", [(int(a), round(float(t[a]), 1)) for a in order[:12]])
    print("fastest:", [(int(a), round(float(t[a]), 1)) for a in order[-8:]])
    m.close()
