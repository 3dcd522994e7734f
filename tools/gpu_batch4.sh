#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
tools/gpu_ab.sh "--steps 6 --warmup 3" bl16 nosort b128 b128r b512
