#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "#### streams"
for k in 1 2 3; do
  RTB_WF_STREAMS=$k timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/st_$k.log 2> gpurun_out/st_$k.err
  echo "streams $k: $(python -c "
import json;d=json.loads(open('gpurun_out/st_$k.log').read().strip().splitlines()[-1]);print('value %.1f ms/step %.2f e2e %.1f'%(d['value'],d['ms_per_step'],d['e2e']['value']))")"
done
echo "#### all configs, 1 GPU (rows-per-step = whole grid)"
for w in c1 c2 c3 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --rows-per-step 100 --no-cpu-baseline > gpurun_out/cfg_$w.log 2> gpurun_out/cfg_$w.err
  echo "$w: $(python -c "
import json;d=json.loads(open('gpurun_out/cfg_$w.log').read().strip().splitlines()[-1]);r=d['roofline'];print('value %.1f Mpaths/s  ms/step %.2f  paths/step %d  e2e %.1f  seg/path %.2f nodes/seg %.2f prims/seg %.2f  frac %.3f'%(d['value'],d['ms_per_step'],d['config']['paths_per_step_per_gpu'],d['e2e']['value'],r['segments_per_path'],r['node_visits_per_segment'],r['prim_tests_per_segment'],r['frac']))" 2>&1 | tail -1)"
done
