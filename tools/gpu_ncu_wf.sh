#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
CMD="python bench.py --steps 1 --warmup 1 --pipeline wavefront --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_wf2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_wf_extend -s 40 -c 1 -o gpurun_out/prof_pool $CMD > gpurun_out/ncu_full_wf.log 2>&1
echo "full rc=$?"
