#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
tools/gpu_ab.sh "--steps 6 --warmup 3 --rows-per-step 1" default thr20 thr24 thr32 bl12 bl20
