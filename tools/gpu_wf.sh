#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for cfg in "1 4194304 0" "2 4194304 0" "2 8388608 0" "2 8388608 4" "2 8388608 5" "3 8388608 0" "4 8388608 0" "2 4194304 4"; do set -- $cfg; echo "== streams $1 cap $2 extend_blocks $3"; export RTB_WF_STREAMS=$1 RTB_WF_CAPACITY=$2; if [ "$3" != "0" ]; then export RTB_WF_EXTEND_BLOCKS=$3; else unset RTB_WF_EXTEND_BLOCKS; fi; timeout 600 python bench.py --steps 5 --warmup 3 --pipeline wavefront --no-cpu-baseline > gpurun_out/b.log 2> gpurun_out/b.err; python -c "
import json
d=json.loads(open('gpurun_out/b.log').read().strip().splitlines()[-1]); print({k:round(d[k],2) for k in ('value','ms_per_step','gpu_launches')}, round(d['e2e']['value'],1))"; tail -1 gpurun_out/b.err; done
