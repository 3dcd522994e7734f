#!/bin/bash
# round-2 evidence for profiles/: launch list of the bench command + full captures of the hot kernels of the FINAL binary
mkdir -p gpurun_out
export PYTHONPATH=$PWD
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs"
timeout 600 $CMD > gpurun_out/plain_a.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_a.log 2>&1
echo "launch list rc=$?"
timeout 600 $CMD > gpurun_out/plain_b.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_wf_ -s 700 -c 5 -f -o gpurun_out/r02_iteration $CMD > gpurun_out/ncu_b.log 2>&1
echo "full iteration rc=$?"
timeout 600 $CMD --option 2=1 > gpurun_out/plain_c.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_wf_extend<' -s 140 -c 1 -f -o gpurun_out/r02_extend_exact_arm $CMD --option 2=1 > gpurun_out/ncu_c.log 2>&1
echo "exact arm rc=$?"
timeout 600 $CMD --flags 0x400 > gpurun_out/plain_d.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_wf_extend<' -s 140 -c 1 -f -o gpurun_out/r02_extend_sixleaf_arm $CMD --flags 0x400 > gpurun_out/ncu_d.log 2>&1
echo "six-leaf arm rc=$?"
for f in a b c d; do tail -1 gpurun_out/plain_$f.log | cut -c1-160; done
cp surely_raytracing_b200/librtb200.so gpurun_out/librtb200_profiled.so
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches.csv
