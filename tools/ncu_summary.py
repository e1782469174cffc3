#!/usr/bin/env python
"""Summarise ncu outputs into the small text files kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r01.csv            # per-kernel share of a step
  python tools/ncu_summary.py kernel   gpurun_out/prof_vr_r01.ncu-rep         # key metrics of a --set full capture
"""
import collections
import csv
import subprocess
import sys

KEY_METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "smsp__inst_executed.sum",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        a = agg.setdefault(r[kn].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", "")) * scale[r[mu]]
    total = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("# total %.3f ms over %d launches" % (total, sum(a[0] for a in agg.values())))
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-64s launches=%4d  %10.3f ms  %5.1f%%" % (k[:64], n, t, 100 * t / total))


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    print("# ncu --set full --clock-control none; one column per captured launch")
    print("%-84s %-12s %s" % ("kernel", "", " | ".join(r[name_col].split("(")[0][-40:] for r in rows[2:])))
    for m in KEY_METRICS:
        for i, h in enumerate(hdr):
            if h != m and not h.endswith("." + m):
                continue
            print("%-84s %-12s %s" % (m, units[i], " | ".join(r[i] for r in rows[2:])))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
