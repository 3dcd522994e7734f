#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== stage profile"; RTB_WF_PROFILE=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_prof.log 2> gpurun_out/bench_prof.err; grep "rtb wavefront" gpurun_out/bench_prof.err | sed -n '2,2p'
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b.log 2> gpurun_out/b.err; python -c "
import json
d=json.loads(open('gpurun_out/b.log').read().strip().splitlines()[-1]); print({k:round(d[k],2) for k in ('value','ms_per_step','gpu_launches')}, round(d['e2e']['value'],1))"; tail -1 gpurun_out/b.err
