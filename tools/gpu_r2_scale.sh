#!/bin/bash
# one box, up to 8 GPUs: rtb_render_multi test, weak scaling N = 1, 2, 4, 8 back to back, strong scaling (the whole 6.4 G-path
# c4 job, image checked) at the largest N
mkdir -p gpurun_out
export PYTHONPATH=$PWD
MAXN=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "render_multi" -s > gpurun_out/pytest_multi_g$MAXN.log 2>&1; echo "pytest multi rc=$?"; tail -4 gpurun_out/pytest_multi_g$MAXN.log
summ() {
  python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   n_gpus %d value %.1f  ms/step %.2f  e2e %.1f  scaling %s  reduce_ms %.3f  image_check %s"%(d["n_gpus"],d["value"],d["ms_per_step"],d["e2e"]["value"],d["scaling"],d["config"]["reduce_ms"],json.dumps(d.get("image_check"))[:400]))
except Exception as e: print("   no json",e)
P
}
for N in 1 2 4 8; do
  [ $N -gt $MAXN ] && continue
  if [ $N -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"; fi
  timeout 900 $L bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/scale_weak_g$N.log 2> gpurun_out/scale_weak_g$N.err
  echo "== weak N=$N rc=$?"; summ gpurun_out/scale_weak_g$N.log
done
N=$MAXN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-cpu-baseline --no-configs > gpurun_out/scale_strong_g$N.log 2> gpurun_out/scale_strong_g$N.err
echo "== strong N=$N rc=$?"; summ gpurun_out/scale_strong_g$N.log
