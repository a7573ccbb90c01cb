"""Where does the fused tick's small kernel spend its time? Statistics build (libpmaf_stats.so, -DPMAF_FAST_STATS):
tick_kernel stamps %globaltimer at its phase boundaries.

    make -C predictive-multi-agent-framework_b200/csrc OUT=../libpmaf_stats.so EXTRA=-DPMAF_FAST_STATS
    python tools/tick_phases.py [c2|c5] [ticks]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pmaf_b200  # noqa: E402,F401
from pmaf_b200 import loop, planner, scenarios  # noqa: E402

if __name__ == "__main__":
    planner.LIB_PATH = os.path.join(os.path.dirname(planner.LIB_PATH), "libpmaf_stats.so")
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    sc = getattr(scenarios, name)()
    m = planner.CfManager(0)
    feed = loop.ObstacleFeed(sc)
    loop.plan_begin(m, sc)
    rows = []
    for t in range(ticks):
        m.flush_l2()
        m.stop_prediction()
        m.timer_start()
        m.tick(feed.pos, feed.vel, feed.rad, sc.delta_t, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist, sc.k_workspace,
               sc.ws_limits)
        ms = m.timer_stop()
        feed.step()
        c = m.counters()
        st = list(m.fast_stats().values())  # step_counter[4..15]; stamps are [8..11] -> indices 4..7
        s = np.array(st[4:8], dtype=np.float64)
        rows.append([ms * 1e3, c["last_rollout_ms"] * 1e3, (s[1] - s[0]) / 1e3, (s[2] - s[1]) / 1e3, (s[3] - s[2]) / 1e3])
    r = np.array(rows[3:])
    print(f"{sc.name}: median over {len(r)} ticks [us]: tick {np.median(r[:, 0]):.1f}  rollout (events) {np.median(r[:, 1]):.1f}  "
          f"tick_kernel: prefetch+evaluate {np.median(r[:, 2]):.1f}  real step / image {np.median(r[:, 3]):.1f}  "
          f"known words {np.median(r[:, 4]):.1f}")
    m.close()
