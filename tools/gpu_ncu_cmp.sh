#!/bin/bash
# full ncu capture of one steady-state k_wf_extend + k_wf_shade launch, BVH2 vs BVH4 (same library)
mkdir -p gpurun_out
export PYTHONPATH=$PWD
CMD="python bench.py --steps 1 --warmup 1 --rows-per-step 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_cmp.log 2>&1 || { echo "plain run failed"; exit 1; }
for v in 0 1; do
  RTB_BVH4=$v timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_wf_ -s 120 -c 4 -f -o gpurun_out/prof_bvh4_$v $CMD > gpurun_out/ncu_cmp_$v.log 2>&1
  echo "bvh4=$v rc=$?"
done
ls -la gpurun_out/*.ncu-rep
