"""SASS evidence of the in-tree libpmaf.so, per kernel: instruction totals and the counts of the Blackwell / Hopper-era
mnemonics the design relies on (UBLKCP = TMA bulk copy, SYNCS = mbarrier, CREDUX / REDUX = warp-wide integer
reductions, VOTE / SHFL / MATCH = collectives, ACQBULK / griddepcontrol as emitted), plus `ptxas -v` resources.

    python tools/sass_counts.py > profiles/r02_sass_counts.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "predictive-multi-agent-framework_b200", "libpmaf.so")
WATCH = ["UBLKCP", "SYNCS", "CREDUX", "REDUX", "VOTE", "VOTEU", "SHFL", "MATCH", "WARPSYNC", "BAR", "MUFU", "DFMA", "DADD", "DMUL",
         "BSSY", "BSYNC", "BRA", "CALL", "LDL", "STL", "ACQBULK", "PREEXIT", "CCTL"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except OSError:
        return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            counts[cur][m.group(1)] += 1
            counts[cur]["_total"] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a), static instruction counts per kernel")
    for fn, c in counts.items():
        if c["_total"] == 0:
            continue
        hits = "  ".join(f"{k}={c[k]}" for k in WATCH if c[k])
        print(f"{demangle(fn)}\n    instructions={c['_total']}  {hits}")
    print("\n# ptxas -v (make -C csrc ptxas-info): registers / spills per kernel")
    out = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "predictive-multi-agent-framework_b200", "csrc"), "ptxas-info"],
                         capture_output=True, text=True).stdout
    fn = None
    for line in out.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            fn = demangle(m.group(1))
        elif "registers" in line and fn:
            print(f"{fn}\n    {line.strip().replace('ptxas info    : ', '')}")
        elif "spill" in line and fn:
            print(f"    {line.strip()}")


if __name__ == "__main__":
    main()
