import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import pmaf_b200
from pmaf_b200 import loop, planner, scenarios
sc = scenarios.c2()
m = planner.CfManager(0)
feed = loop.ObstacleFeed(sc); loop.plan_begin(m, sc)
for t in range(12):
    m.flush_l2(); loop.control_tick(m, sc, feed); feed.step(); m.stop_prediction()
s = m.get_agent_summaries()
t = s["pred_time_ns"]; 
print("rollout ms", m.counters()["last_rollout_ms"])
print("per-agent us: min %.1f med %.1f p90 %.1f max %.1f" % (t.min()/1e3, np.median(t)/1e3, np.percentile(t,90)/1e3, t.max()/1e3))
order = np.argsort(-t)
print("slowest:", [(int(a), round(float(t[a])/1e3,1)) for a in order[:12]])
print("fastest:", [(int(a), round(float(t[a])/1e3,1)) for a in order[-8:]])
print("first 8 agents:", [round(float(x)/1e3,1) for x in t[:8]])
print("hist:", np.histogram(t/1e3, bins=10))
m.close()
