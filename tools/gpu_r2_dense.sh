#!/bin/bash
# dense leaf phase variant: parity subset through the variant library, then A/B
mkdir -p gpurun_out
export PYTHONPATH=$PWD
RTB200_LIB=$PWD/surely_raytracing_b200/variants/librtb200_$1.so timeout 900 python -m pytest tests -m gpu -q -x -k "candidate or first_hit or exact_image or tree or box_leaves or reproducible" > gpurun_out/pytest_variant.log 2>&1; echo "variant pytest rc=$?"; grep -E "^E  |Error|assert" gpurun_out/pytest_variant.log | head -8; tail -2 gpurun_out/pytest_variant.log
bash tools/gpu_r2_ab.sh "" default "$@" default
