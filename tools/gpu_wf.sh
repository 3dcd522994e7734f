#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for v in k3 k4 k10 k16; do
  export RTB200_LIB=$PWD/surely_raytracing_b200/librtb200_$v.so
  echo "== $v"; RTB_WF_PROFILE=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_prof.log 2> gpurun_out/bench_prof.err; grep "rtb wavefront" gpurun_out/bench_prof.err | sed -n '2,2p'
done
