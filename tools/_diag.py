import sys, os, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pmaf_b200
from pmaf_b200 import loop, scenarios, planner
sc = scenarios.c2()
for sampler in (False, True, False, True):
    p = None
    if sampler:
        p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.DEVNULL)
        time.sleep(0.3)
    mgr = planner.CfManager(0)
    feed = loop.ObstacleFeed(sc)
    loop.plan_begin(mgr, sc)
    for _ in range(5):
        mgr.tick(feed.pos, feed.vel, feed.rad, sc.delta_t, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist, sc.k_workspace, sc.ws_limits)
    mgr.stop_prediction()
    mgr.set_upload_dedup(False)
    ts = []
    for rep in range(6):
        s, _, _, _ = mgr.dry_run(50, feed.pos, feed.vel, feed.rad, 0, sc.delta_t, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist, sc.k_workspace, sc.ws_limits, feed_frequency=100.0, wait_rollout=True, flush_l2=True)
        ts.append(1e3 * s / 50)
    dev = []
    for rep in range(6):
        tot = 0.0
        for _ in range(50):
            mgr.flush_l2(); mgr.stop_prediction(); mgr.timer_start()
            mgr.tick(feed.pos, feed.vel, feed.rad, sc.delta_t, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist, sc.k_workspace, sc.ws_limits)
            tot += mgr.timer_stop()
        dev.append(tot / 50)
    print("sampler", sampler, "e2e ms/tick", [round(t, 4) for t in ts], "device ms/tick", [round(t, 4) for t in dev], flush=True)
    mgr.close()
    if p:
        p.terminate(); p.wait()
