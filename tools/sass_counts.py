#!/usr/bin/env python
"""Static SASS instruction counts per kernel of libxmlb200.so (evidence that the tensor-core kernels really are
tcgen05 / TMA code):  python tools/sass_counts.py > profiles/r02_sass_tc_kernels.txt"""
import collections
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "tvretrieval_b200", "libxmlb200.so")
COLS = [("UTC*MMA", r"\bUTC\w*MMA\b"), ("UTMALDG", r"\bUTMALDG\b"), ("UBLKCP", r"\bUBLKCP\b"), ("LDTM", r"\bLDTM\b"),
        ("UTCBAR", r"\bUTCBAR\b"), ("ELECT", r"\bELECT\b"), ("SYNCS", r"\bSYNCS\b"), ("LDGSTS", r"\bLDGSTS\b"),
        ("HMMA", r"\bHMMA\b"), ("FFMA", r"\bFFMA\b"), ("STG", r"\bSTG\b")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name).split("(")[0].replace("void ", "")
            cur = counts.setdefault(name, collections.Counter())
            continue
        if cur is None or "/*" not in line:
            continue
        for col, pat in COLS:
            if re.search(pat, line):
                cur[col] += 1
    print("# cuobjdump -sass tvretrieval_b200/libxmlb200.so (sm_100a): static instruction counts per kernel")
    print("# UTC*MMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk (plain bulk copy), LDTM ="
          " tcgen05.ld,\n# UTCBAR = tcgen05.commit, ELECT = elect.sync (also the compiler's elect-and-retry loops around"
          " single-thread tcgen05 / TMA\n# issue), SYNCS = mbarrier ops, LDGSTS = cp.async, HMMA = legacy mma.sync"
          " (absent everywhere)\n")
    print("%-44s" % "kernel" + "".join("%9s" % c for c, _ in COLS))
    order = sorted(counts.items(), key=lambda kv: (-kv[1]["UTC*MMA"], kv[0]))
    for name, c in order:
        print("%-44s" % name[:44] + "".join("%9d" % c[col] for col, _ in COLS))


if __name__ == "__main__":
    main()
