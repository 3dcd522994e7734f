#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for cap in 1048576 4194304; do echo "== stage profile cap $cap"; RTB_WF_PROFILE=1 RTB_WF_CAPACITY=$cap timeout 600 python bench.py --steps 2 --warmup 1 --pipeline wavefront --no-cpu-baseline > gpurun_out/bench_prof_$cap.log 2> gpurun_out/bench_prof_$cap.err; grep "rtb wavefront" gpurun_out/bench_prof_$cap.err | sed -n '2,2p'; done
for cap in 2097152 4194304; do echo "== bench wavefront cap $cap"; RTB_WF_CAPACITY=$cap timeout 600 python bench.py --steps 5 --warmup 3 --pipeline wavefront --no-cpu-baseline > gpurun_out/bench_wf_$cap.log 2> gpurun_out/bench_wf_$cap.err; echo "rc=$?"; python -c "
import json,sys
l=open('gpurun_out/bench_wf_$cap.log').read().strip().splitlines()[-1]; d=json.loads(l)
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
"; tail -2 gpurun_out/bench_wf_$cap.err; done
