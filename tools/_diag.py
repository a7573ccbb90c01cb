import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pmaf_b200
from pmaf_b200 import loop, scenarios, planner
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
sc = scenarios.c2()
mgr = planner.CfManager(0)
mgr.set_rollout_timing(False)
class S(threading.Thread):
    def __init__(self, period, power):
        super().__init__(daemon=True); self.period=period; self.power=power; self.stop_=False; self.n=0; self.clk=[]; self.tq=[]
    def run(self):
        while not self.stop_:
            t0=time.perf_counter()
            self.clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml,'nvmlDeviceGetCurrentClocksEventReasons') else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            if self.power: pynvml.nvmlDeviceGetPowerUsage(h)
            self.tq.append(time.perf_counter()-t0)
            self.n+=1
            time.sleep(self.period)
for name, period, power in (("none", 0, False), ("nvml/20ms", 0.02, False), ("nvml/20ms+power", 0.02, True), ("nvml/5ms", 0.005, False), ("none", 0, False)):
    s = None
    if period:
        s = S(period, power); s.start()
    feed = loop.ObstacleFeed(sc)
    loop.plan_begin(mgr, sc)
    ts = []
    for i in range(3000):
        if i % 100 == 0:
            loop.plan_begin(mgr, sc)
        mgr.timer_start()
        mgr.tick(feed.pos, feed.vel, feed.rad, sc.delta_t, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist, sc.k_workspace, sc.ws_limits)
        ts.append(mgr.timer_stop())
    # e2e style
    mgr.set_upload_dedup(False)
    tt = []
    for rep in range(10):
        loop.plan_begin(mgr, sc)
        tick_s = []
        mgr.dry_run(100, feed.pos, feed.vel, feed.rad, 0, sc.delta_t, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist, sc.k_workspace, sc.ws_limits, wait_rollout=True, flush_l2=True, tick_times=tick_s)
        tt += tick_s
    mgr.set_upload_dedup(True)
    ts = np.array(ts); tt = 1e3*np.array(tt)
    med = np.median(ts); st = ts[ts > 1.5*med]; med2 = np.median(tt); st2 = tt[tt > 1.5*med2]
    extra = f" samples {s.n} query ms median {1e3*np.median(s.tq):.3f} max {1e3*np.max(s.tq):.3f} clk {np.median(s.clk)}" if s else ""
    print(f"{name:16s} device: median {med:.4f} stalls {len(st)} worst {ts.max():.3f} | e2e: median {med2:.4f} stalls {len(st2)} worst {tt.max():.3f} ms{extra}", flush=True)
    if s:
        s.stop_ = True; s.join()
mgr.close()
