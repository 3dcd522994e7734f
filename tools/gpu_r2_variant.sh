#!/bin/bash
# a compile-time variant (tools/variants.py): parity subset through the variant library, then A/B on c4 and c5
mkdir -p gpurun_out
export PYTHONPATH=$PWD
RTB200_LIB=$PWD/surely_raytracing_b200/variants/librtb200_$1.so timeout 900 python -m pytest tests -m gpu -q -x -k "candidate or first_hit or exact_image or tree or box_leaves or reproducible" > gpurun_out/pytest_variant.log 2>&1; echo "variant pytest rc=$?"; grep -E "^E  |Error|assert" gpurun_out/pytest_variant.log | head -8; tail -2 gpurun_out/pytest_variant.log
bash tools/gpu_r2_ab.sh "" default "$@" default
bash tools/gpu_r2_ab.sh "--workload c5" default "$@"
