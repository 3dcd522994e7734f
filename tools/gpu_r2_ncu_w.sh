#!/bin/bash
# one full capture of a steady-state launch of one kernel on a workload: tools/gpu_r2_ncu_w.sh <workload> <kernel-regex> <skip> <out-name>
mkdir -p gpurun_out
export PYTHONPATH=$PWD
CMD="python bench.py --workload $1 --steps 1 --warmup 1 --no-cpu-baseline --no-configs"
timeout 600 $CMD > gpurun_out/plain_w.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$2" -s $3 -c 1 -f -o gpurun_out/$4 $CMD > gpurun_out/ncu_w.log 2>&1
echo "capture rc=$?"; ls -la gpurun_out/$4.ncu-rep; tail -1 gpurun_out/plain_w.log | cut -c1-200
