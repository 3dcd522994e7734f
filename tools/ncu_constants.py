#!/usr/bin/env python
"""bench.py's NCU block from an ncu launch list of the bench command (tools/gpu_r2_profile.sh) and the JSON line of the
plain run of the same command: executed thread-instructions per segment of the two hot kernels, their lanes per warp
instruction, DRAM bytes per segment over every k_wf_* launch.

usage: tools/ncu_constants.py <launches.csv> <plain bench log>"""
import collections
import csv
import json
import sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
line = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("void rtb::", "").replace("void ", "").replace("rtb::", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"].lower()
    if r["Metric Name"].startswith("dram__bytes"):
        v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
    if r["Metric Name"].startswith("gpu__time"):
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)   # -> us
    agg[k][r["Metric Name"]] += v
    if r["Metric Name"] == "gpu__time_duration.sum":
        cnt[k] += 1
roof = line["roofline"]
segs_per_pass = line["config"]["paths_per_step"] * roof["segments_per_path"]
ext = next(k for k in agg if k.startswith("k_wf_extend<0"))
shd = next(k for k in agg if k.startswith("k_wf_shade<0"))
# full-size passes with the non-counting kernels = k_wf_init launches minus the counted pass (the <1, ...> instantiations)
counted = 1 if any(k.startswith("k_wf_extend<1") for k in agg) else 0
passes = cnt["k_wf_init"] - counted
print(f"passes of {line['config']['paths_per_step']} paths in the list: {passes} ({cnt[ext]} iterations); segments per pass {segs_per_pass:.4g}")
tot_time = sum(v["gpu__time_duration.sum"] for k, v in agg.items() if k.startswith("k_wf_"))
out = {"thread_instr_per_segment": {}, "lanes": {}}
for name, k in (("k_wf_extend", ext), ("k_wf_shade", shd)):
    ti = agg[k]["smsp__thread_inst_executed.sum"]
    wi = agg[k]["smsp__inst_executed.sum"]
    out["thread_instr_per_segment"][name] = round(ti / (passes * segs_per_pass), 1)
    out["lanes"][name] = round(ti / wi, 2)
    print(f"{name}: thread-instr/segment {ti / (passes * segs_per_pass):.1f}  lanes {ti / wi:.2f}  share of k_wf_* time {agg[k]['gpu__time_duration.sum'] / tot_time:.3f}")
dram = sum(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"] for k, v in agg.items() if k.startswith("k_wf_") and not k.startswith(("k_wf_extend<1", "k_wf_shade<1")))
out["dram_bytes_per_segment_c4"] = round(dram / (passes * segs_per_pass), 1)
print(f"DRAM bytes per segment over all k_wf_* launches: {dram / (passes * segs_per_pass):.1f}  ({dram / passes / 1e9:.1f} GB per step)")
print(json.dumps(out))
