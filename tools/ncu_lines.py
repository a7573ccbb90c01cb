"""Per-source-line instruction counts of an `ncu --set full --import-source on` report (cuda,sass view).

    python tools/ncu_lines.py report.ncu-rep agent_steps_per_launch [top]

Prints, per file:line, the warp instructions executed per agent-step (inlined code is attributed to the
line it came from), the FP64 share, the average active threads and the share of the stall samples.
The correlated view lists a SASS instruction under every source line it is attributed to (call site and
inlined callee), so the per-line counts add up to ~20 % more than the kernel's instruction total: read them as a
ranking, and take totals from tools/summarize_ncu.py.
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, steps = sys.argv[1], float(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    per_line = collections.defaultdict(lambda: [0, 0, 0, 0, ""])  # inst, fp64 inst, thread inst, samples, text
    fname, hdr, cur = "", None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_inst, i_thr, i_smp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) <= i_thr:
            continue
        if r[0]:  # a source line
            cur = (fname, int(r[0]))
            per_line[cur][4] = r[1].strip()
            continue
        if cur is None or r[2] in ("...", "-"):
            continue
        try:
            n, t, s = int(r[i_inst]), int(r[i_thr]), int(r[i_smp])
        except ValueError:
            continue
        e = per_line[cur]
        e[0] += n
        e[2] += t
        e[3] += s
        if re.match(r"\s*(?:@!?U?P\d+\s+)?(DADD|DMUL|DFMA|DSETP|MUFU)", r[3]):
            e[1] += n
    tot = sum(e[0] for e in per_line.values())
    smp = sum(e[3] for e in per_line.values())
    print(f"total warp instructions / agent-step: {tot / steps:.1f}; samples {smp}")
    for (f, ln), e in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        if e[0] == 0:
            continue
        print(f"{e[0] / steps:8.1f} inst  fp64 {e[1] / steps:7.1f}  thr {e[2] / max(e[0], 1):5.1f}  smp {100 * e[3] / max(smp, 1):5.2f}%  {f}:{ln}  {e[4][:90]}")


if __name__ == "__main__":
    main()
