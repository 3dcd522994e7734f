#!/bin/bash
# round-2 first GPU pass: smoke, the gpu tests, A/B of the extend arms (candidates vs exact leaves, occupancy variants)
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -2
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
ARGS="--steps 3 --warmup 1 --no-cpu-baseline --no-configs"
run() {  # name, lib-variant-or-default, extra args
  if [ "$2" = "default" ]; then unset RTB200_LIB; else export RTB200_LIB=$PWD/surely_raytracing_b200/variants/librtb200_$2.so; fi
  timeout 600 python bench.py $ARGS $3 > gpurun_out/ab_$1.log 2> gpurun_out/ab_$1.err
  echo "== $1 rc=$? $(grep step_ms gpurun_out/ab_$1.err | cut -c1-100)"
  python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    r=d["roofline"]; k=r["kernels"]
    print("   value %.1f  ms/step %.2f  e2e %.1f  frac %.3f  nodes/seg %.2f prims/seg %.3f exact/seg %.3f ovf/seg %.5f  extend share %.3f shade share %.3f"%(
        d["value"],d["ms_per_step"],d["e2e"]["value"],r["frac"],r["node_visits_per_segment"],r["prim_tests_per_segment"],r["exact_tests_per_segment"],
        r["overflow_rays_per_segment"],k["k_wf_extend"]["share_of_step"],k["k_wf_shade"]["share_of_step"]))
except Exception as e: print("   no json",e); print(open(f"gpurun_out/ab_{sys.argv[1]}.err").read()[-1500:])
P
}
run cand default ""
run exact default "--option 2=1"
run smem default "--option 3=1"
for v in "$@"; do run $v $v ""; done
