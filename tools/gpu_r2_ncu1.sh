#!/bin/bash
# one full capture of one k_wf_extend launch of a variant library: tools/gpu_r2_ncu1.sh <variant> <out-name>
mkdir -p gpurun_out
export PYTHONPATH=$PWD
export RTB200_LIB=$PWD/surely_raytracing_b200/variants/librtb200_$1.so
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs"
timeout 600 $CMD > gpurun_out/plain_v.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_wf_extend<' -s 140 -c 1 -f -o gpurun_out/$2 $CMD > gpurun_out/ncu_v.log 2>&1
echo "capture rc=$?"; ls -la gpurun_out/$2.ncu-rep
