#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
CMD="python bench.py --steps 1 --warmup 1 --rows-per-step 1 --no-cpu-baseline"
RTB_BVH4=0 RTB_QNODES=1 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_wf_extend -s 60 -c 1 -f -o gpurun_out/prof_q $CMD > gpurun_out/ncu_q.log 2>&1
echo "q rc=$?"
