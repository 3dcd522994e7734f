#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi -L | wc -l
for n in "$@"; do
  echo "== bench --gpus $n"
  if [ "$n" = "1" ]; then timeout 900 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g$n.log 2> gpurun_out/bench_g$n.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/bench_g$n.log 2> gpurun_out/bench_g$n.err; fi
  echo "rc=$?"; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_g$n.log').read().strip().splitlines() if l.startswith('{')][-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, 'e2e', d['e2e']['value'], 'reduce_ms', d['config']['reduce_ms'], d['clocks'])"; grep -E "step_ms|e2e" gpurun_out/bench_g$n.err | tail -2
done
