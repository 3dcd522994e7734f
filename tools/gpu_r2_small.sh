#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^E  |Error|assert" gpurun_out/pytest_gpu.log | head -12; tail -2 gpurun_out/pytest_gpu.log
bash tools/gpu_r2_ab.sh "--workload c1" default
bash tools/gpu_r2_ab.sh "" default
