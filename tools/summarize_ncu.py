"""Summarise an `ncu --set full` report of the rollout kernel into the JSON kept under profiles/.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep "<description>" agent_steps_per_launch > profiles/NAME.json
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        # lane utilisation: threads per executed warp instruction, all / with their predicate on
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep, desc, steps = sys.argv[1], sys.argv[2], float(sys.argv[3])
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    metrics = {h: {"unit": u, "value": v} for h, u, v in zip(hdr, units, vals) if h in KEEP}
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    h2 = rows[1]
    ia, ie = h2.index("Source"), h2.index("Instructions Executed")
    ops = collections.Counter()
    for r in rows[2:]:
        if len(r) <= ie:
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
        try:
            ops[m.group(1).split(".")[0] if m else "?"] += int(r[ie])
        except ValueError:
            pass
    tot = sum(ops.values())
    # warp-wide integer minima compile to CREDUX (redux.sync), the bulk copy to UBLKCP, its mbarrier to SYNCS
    tma = {k: ops.get(k, 0) for k in ("UBLKCP", "SYNCS", "CREDUX", "REDUX", "MUFU", "SHFL", "BAR", "WARPSYNC")}
    def f(name):
        return float(metrics[name]["value"].replace(",", ""))
    out = {"kernel": desc, "metrics": metrics,
           "dram_bytes_per_launch": None, "warp_instructions_per_agent_step": tot / steps,
           "opcode_mix_pct": {k: round(100 * v / tot, 2) for k, v in ops.most_common(18)},
           "blackwell_evidence_instruction_counts": tma}
    try:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        out["dram_bytes_per_launch"] = sum(f(k) * scale[metrics[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except Exception:
        pass
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
