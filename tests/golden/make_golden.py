"""Regenerate tests/golden/*.npz from the REFERENCE build (oracle/_ref/libcfref.so = the
reference's unmodified cf_agent.cpp / cf_manager.cpp, see oracle/Makefile). Run in the dev
container, where /root/reference exists:

    python tests/golden/make_golden.py

Every file holds the outputs of one named case of pmaf_b200.cases (inputs are re-derived from
the case's seeds; `sha_inputs` guards against generator drift)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import pmaf_b200  # noqa: E402,F401
from pmaf_b200 import cases, scenarios  # noqa: E402
from oracle import cpu_planners  # noqa: E402


def input_digest(sc):
    h = hashlib.sha256()
    for a in (sc.goal, sc.start, sc.obs_pos, sc.obs_vel, sc.obs_rad, sc.random_vecs(), *sc.gains().values()):
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


def main():
    cpu_planners.build("ref")
    out_dir = os.path.dirname(os.path.abspath(__file__))
    total = 0
    for name, case in cases.all_cases().items():
        p = cpu_planners.RefPlanner(pooled=True)
        rec = case(p)
        p.close()
        rec["sha_inputs"] = np.array(input_digest(case.scenario))
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **rec)
        total += os.path.getsize(path)
        print(f"{name:28s} {os.path.getsize(path) / 1024:8.1f} KiB  best={rec['best'][:8]}")
    # the reference's own task files (all nine config/tasks/*.yaml; dual_arms_static1 is the anchor above)
    tasks = "/root/reference/src/bimanual_planning_ros/config/tasks"
    starts = {"dual_arms": (-0.6, 0.0, 0.65), "sim_kobo": (0.25, -0.35, 0.45)}
    for fn in sorted(os.listdir(tasks)) if os.path.isdir(tasks) else []:
        start = starts["sim_kobo" if fn.startswith("sim_kobo") else "dual_arms"]
        sc = scenarios.from_task_yaml(os.path.join(tasks, fn), start=start, seed=11)
        sc = sc.with_(max_prediction_steps=min(sc.max_prediction_steps, 600))  # keep the files small
        ticks = 60
        p = cpu_planners.RefPlanner(pooled=True)
        rec = cases.closed_loop(sc, ticks)(p)
        p.close()
        rec.update(scenarios.to_arrays(sc))
        rec["in_ticks"] = np.array(ticks)
        rec["sha_inputs"] = np.array(input_digest(sc))
        path = os.path.join(out_dir, "task_" + sc.name + ".npz")
        np.savez_compressed(path, **rec)
        total += os.path.getsize(path)
        print(f"task_{sc.name:23s} {os.path.getsize(path) / 1024:8.1f} KiB  best={np.unique(rec['best'])}")
    print(f"total {total / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
