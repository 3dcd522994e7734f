#!/bin/bash
# round-end check without the CPU reference arm: smoke, gpu tests, default bench
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --cpu-seconds 4 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-260 gpurun_out/bench.log; tail -2 gpurun_out/bench.err
