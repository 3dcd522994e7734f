#!/bin/bash
# round-1 evidence for profiles/: launch list of the bench command + full captures of the two hot kernels
mkdir -p gpurun_out
export PYTHONPATH=$PWD
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_a.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_a.log 2>&1
echo "launch list rc=$?"
timeout 600 $CMD > gpurun_out/plain_b.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_wf_ -s 200 -c 4 -f -o gpurun_out/prof_final2 $CMD > gpurun_out/ncu_b.log 2>&1
echo "full rc=$?"
RTB_WF_PROFILE=1 timeout 600 $CMD > gpurun_out/stage.log 2> gpurun_out/stage.err; grep "rtb wavefront" gpurun_out/stage.err
tail -1 gpurun_out/plain_a.log | cut -c1-300
ls -la gpurun_out | grep -E "final"
