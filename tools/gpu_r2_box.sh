#!/bin/bash
# box leaves A/B: default build (make_box = one leaf) against RTB_FLAG_NO_BOX_LEAVES (0x400), tests first
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^E  |Error|assert" gpurun_out/pytest_gpu.log | head -12; tail -3 gpurun_out/pytest_gpu.log
bash tools/gpu_r2_ab.sh "" default
bash tools/gpu_r2_ab.sh "--flags 0x400" default
bash tools/gpu_r2_ab.sh "" opt:2=1
