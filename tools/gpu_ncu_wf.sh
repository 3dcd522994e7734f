#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
CMD="python bench.py --steps 1 --warmup 1 --pipeline wavefront --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_wf.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 800 -c 300 --csv --log-file gpurun_out/launches_wf.csv $CMD > gpurun_out/ncu_launches_wf.log 2>&1
echo "launch list rc=$?"
timeout 600 $CMD > gpurun_out/plain_wf2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_wf_ -s 500 -c 5 -o gpurun_out/prof_wf $CMD > gpurun_out/ncu_full_wf.log 2>&1
echo "full rc=$?"; ls -la gpurun_out | grep -E "prof_wf|launches_wf"
