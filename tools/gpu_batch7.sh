#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
export RTB_BVH4=0
for q in 0 1 0 1; do
  export RTB_QNODES=$q
  echo "#### RTB_QNODES=$q"
  tools/gpu_ab.sh "--steps 6 --warmup 3 --rows-per-step 1" default
done
