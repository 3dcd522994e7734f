#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for v in lohi new; do
  if [ "$v" = "lohi" ]; then export RTB200_LIB=$PWD/surely_raytracing_b200/librtb200_lohi.so; else unset RTB200_LIB; fi
  echo "== $v"; timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/b_$v.log 2> gpurun_out/b_$v.err; grep step_ms gpurun_out/b_$v.err
  RTB_WF_PROFILE=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_prof.log 2> gpurun_out/bench_prof.err; grep "rtb wavefront" gpurun_out/bench_prof.err | sed -n '2,2p'
done
