#!/bin/bash
# A/B of library variants on the bench command: tools/gpu_r2_ab.sh "<extra bench args>" variant...   ("default" = in-tree library; "opt:<id>=<v>" = in-tree library with an RTB_OPT)
mkdir -p gpurun_out
export PYTHONPATH=$PWD
ARGS="--steps 3 --warmup 1 --no-cpu-baseline --no-configs $1"; shift
for v in "$@"; do
  extra=""
  case "$v" in
    default) unset RTB200_LIB;;
    opt:*) unset RTB200_LIB; extra="--option ${v#opt:}";;
    *) export RTB200_LIB=$PWD/surely_raytracing_b200/variants/librtb200_$v.so;;
  esac
  n=${v//[:=]/_}
  timeout 600 python bench.py $ARGS $extra > gpurun_out/ab_$n.log 2> gpurun_out/ab_$n.err
  echo "== $v rc=$? $(grep step_ms gpurun_out/ab_$n.err | cut -c1-100)"
  python - "$n" <<'P'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    r=d["roofline"]; k=r["kernels"]
    tot=d["ms_per_step"]
    print("   value %.1f  ms/step %.2f  e2e %.1f  frac %.3f | extend %.1f ms  shade %.1f ms  gen %.1f ms | nodes/seg %.2f exact/seg %.3f ovf/seg %.5f"%(
        d["value"],tot,d["e2e"]["value"],r["frac"],k["k_wf_extend"]["share_of_step"]*tot,k["k_wf_shade"]["share_of_step"]*tot,k["k_wf_generate"]["share_of_step"]*tot,
        r["node_visits_per_segment"],r["exact_tests_per_segment"],r["overflow_rays_per_segment"]))
except Exception as e: print("   no json",e); print(open(f"gpurun_out/ab_{sys.argv[1]}.err").read()[-1500:])
P
done
