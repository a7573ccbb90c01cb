"""Where does a call-by-call control tick spend its wall time? (pmaf_dry_run with PMAF_DRY_RUN_PROFILE)
    python tools/e2e_profile.py [c2|c5] [ticks]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pmaf_b200  # noqa: E402,F401
from pmaf_b200 import loop, planner, scenarios  # noqa: E402

if __name__ == "__main__":
    sc = getattr(scenarios, sys.argv[1] if len(sys.argv) > 1 else "c2")()
    ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    m = planner.CfManager(0)
    m.set_upload_dedup(False)
    feed = loop.ObstacleFeed(sc)
    loop.plan_begin(m, sc)
    n_feed = sc.num_obstacles - 1 if feed.active else 0
    args = (feed.pos, feed.vel, feed.rad, n_feed, sc.delta_t, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist, sc.k_workspace,
            sc.ws_limits)
    m.dry_run(20, *args, wait_rollout=True, flush_l2=True)
    prof = [0.0] * 6
    c0 = m.counters()
    sec, *_ = m.dry_run(ticks, *args, wait_rollout=True, flush_l2=True, profile=prof)
    c1 = m.counters()
    names = ["stop_prediction", "evaluate_agents", "move_real_agent", "get_next + reset_agents", "start_prediction",
             "wait for the rollout"]
    print(f"{sc.name}: {1e6 * sec / ticks:.1f} us per tick; rollout kernel {1e3 * (c1['rollout_ms_total'] - c0['rollout_ms_total']) / ticks:.1f} us")
    for n, v in zip(names, prof):
        print(f"  {n:26s} {1e6 * v / ticks:7.1f} us")
    m.close()
