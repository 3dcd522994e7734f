#!/bin/bash
# full gpu tests + A/B list: tools/gpu_r2_box.sh variant...
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^E  |Error|assert" gpurun_out/pytest_gpu.log | head -12; tail -3 gpurun_out/pytest_gpu.log
bash tools/gpu_r2_ab.sh "" "$@"
