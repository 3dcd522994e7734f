#!/usr/bin/env python
"""Print the key counters of every kernel in an .ncu-rep (raw page): time, issue, lanes, stalls, pipes, L1."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("=====", d["Kernel Name"][:60])
        for k in KEYS:
            if k in d:
                print(f"   {k:75s} {d[k]:>16s} {u[k]}")
        st = sorted(((float(v), h) for h, v in d.items() if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and v), reverse=True)
        print("   stalls/issue:", ", ".join(f"{h.split('issue_stalled_')[1].split('_per_')[0]} {v:.2f}" for v, h in st[:8]))


if __name__ == "__main__":
    main()
