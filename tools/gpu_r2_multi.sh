#!/bin/bash
# N-GPU checks: rtb_render_multi (threads + NCCL inside the ABI), torchrun weak / strong scaling
mkdir -p gpurun_out
export PYTHONPATH=$PWD
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "render_multi" -s > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -4 gpurun_out/pytest_multi.log
for mode in weak strong; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${2:-3} --warmup ${3:-1} --scaling $mode --no-cpu-baseline --no-configs > gpurun_out/bench_${mode}_g$N.log 2> gpurun_out/bench_${mode}_g$N.err
  echo "== $mode N=$N rc=$?"; grep -E "step_ms|e2e \(" gpurun_out/bench_${mode}_g$N.err | cut -c1-200
  python - gpurun_out/bench_${mode}_g$N.log <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   value %.1f  ms/step %.2f  e2e %.1f  scaling %s  reduce_ms %.3f  image_check %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["scaling"],d["config"]["reduce_ms"],json.dumps(d.get("image_check"))[:300]))
except Exception as e: print("   no json",e)
P
done
