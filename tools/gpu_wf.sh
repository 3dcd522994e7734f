#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for i in 1 2; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b.log 2> gpurun_out/b.err; python -c "
import json
d=json.loads(open('gpurun_out/b.log').read().strip().splitlines()[-1]); print({k:round(d[k],2) for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1))"; grep e2e gpurun_out/b.err; done
