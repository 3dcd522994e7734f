#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
N=${1:-2}
nvidia-smi -L > gpurun_out/gpus.txt; cat gpurun_out/gpus.txt | head -3
for n in 1 $N; do
  echo "== bench --gpus $n"
  if [ "$n" = "1" ]; then timeout 900 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g$n.log 2> gpurun_out/bench_g$n.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/bench_g$n.log 2> gpurun_out/bench_g$n.err; fi
  echo "rc=$?"; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_g$n.log').read().strip().splitlines() if l.startswith('{')][-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'], d['config']['reduce_ms'])"; tail -2 gpurun_out/bench_g$n.err
done
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-200
