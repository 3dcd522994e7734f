#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo skip tests
for cap in 524288 1048576 2097152 4194304 1048576 2097152; do echo "== bench wavefront cap $cap"; RTB_WF_CAPACITY=$cap timeout 600 python bench.py --steps 5 --warmup 3 --pipeline wavefront --no-cpu-baseline > gpurun_out/bench_wf_$cap.log 2> gpurun_out/bench_wf_$cap.err; echo "rc=$?"; python -c "
import json,sys
l=open('gpurun_out/bench_wf_$cap.log').read().strip().splitlines()[-1]; d=json.loads(l)
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], {k:round(d['roofline'][k],3) for k in ('segments_per_path','node_visits_per_segment','prim_tests_per_segment')})
"; tail -3 gpurun_out/bench_wf_$cap.err; done
