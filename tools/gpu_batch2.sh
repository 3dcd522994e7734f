#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
tools/gpu_ab.sh "--steps 6 --warmup 3" head s00 s10 s01 default bl20 bl24
echo "#### capacity sweep (default lib)"
for rows in 4 10; do for cap in 4194304 8388608 16777216; do
  export RTB_WF_CAPACITY=$cap
  timeout 600 python bench.py --steps 4 --warmup 2 --rows-per-step $rows --no-cpu-baseline > gpurun_out/cap_${rows}_$cap.log 2> gpurun_out/cap_${rows}_$cap.err
  echo "rows $rows cap $cap: $(python -c "
import json;d=json.loads(open('gpurun_out/cap_${rows}_$cap.log').read().strip().splitlines()[-1]);print('value %.1f ms/step %.2f e2e %.1f'%(d['value'],d['ms_per_step'],d['e2e']['value']))")"
done; done
