#!/usr/bin/env python
"""Summarise an ncu launch-list CSV (tools/gpu_r2_profile.sh) per kernel: share of time, lanes, issue."""
import collections
import csv
import sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("void rtb::", "").replace("void ", "")
    agg[k][r["Metric Name"]].append(float(r["Metric Value"].replace(",", "")))
tot = sum(sum(v["gpu__time_duration.sum"]) for v in agg.values())
print(f"{'kernel':28s} {'n':>4s} {'time%':>7s} {'avg_us':>8s} {'lanes':>6s} {'warp-instr':>11s} {'warps%':>6s} {'issue%':>6s} {'dramMB':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
    t = v["gpu__time_duration.sum"]
    n = len(t)
    g = lambda m: sum(v[m]) / n if v.get(m) else float("nan")
    dram = (g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) / (1e6 if g("dram__bytes_read.sum") > 1e5 else 1.0)  # bytes or Mbyte, by ncu version
    print(f"{k:28s} {n:4d} {100 * sum(t) / tot:6.2f}% {sum(t) / n / 1e3:8.1f} {g('smsp__thread_inst_executed_per_inst_executed.ratio'):6.2f} "
          f"{g('smsp__inst_executed.sum'):11.4g} {g('sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} {dram:7.1f}")
