#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest gpu (default lib)"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
tools/gpu_ab.sh "--steps 6 --warmup 3" default e8 pf b64
echo "#### rows/capacity (default lib)"
for rc in "4 0" "10 0" "10 33554432" "25 0" "25 33554432"; do set -- $rc; rows=$1; cap=$2
  if [ "$cap" = "0" ]; then unset RTB_WF_CAPACITY; else export RTB_WF_CAPACITY=$cap; fi
  timeout 600 python bench.py --steps 4 --warmup 2 --rows-per-step $rows --no-cpu-baseline > gpurun_out/cap_${rows}_$cap.log 2> gpurun_out/cap_${rows}_$cap.err
  echo "rows $rows cap $cap: $(python -c "
import json;d=json.loads(open('gpurun_out/cap_${rows}_$cap.log').read().strip().splitlines()[-1]);print('value %.1f ms/step %.2f e2e %.1f'%(d['value'],d['ms_per_step'],d['e2e']['value']))")"
done
