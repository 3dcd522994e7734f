#!/bin/bash
# A/B runs of library variants: tools/gpu_ab.sh "<bench args>" variant1 variant2 ...   ("default" = the in-tree library)
mkdir -p gpurun_out
export PYTHONPATH=$PWD
ARGS="$1"; shift
for v in "$@"; do
  if [ "$v" = "default" ]; then unset RTB200_LIB; else export RTB200_LIB=$PWD/surely_raytracing_b200/variants/librtb200_$v.so; fi
  timeout 600 python bench.py $ARGS --no-cpu-baseline > gpurun_out/ab_$v.log 2> gpurun_out/ab_$v.err
  echo "== $v rc=$? $(grep step_ms gpurun_out/ab_$v.err | cut -c1-120)"
  RTB_WF_PROFILE=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/abp_$v.err; echo "   $(grep 'rtb wavefront' gpurun_out/abp_$v.err | sed -n '2,2p' | sed 's/.*segments [0-9]* //')"
  python - "$v" <<'P'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    print("   value %.1f  ms/step %.2f  e2e %.1f  nodes/seg %.2f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["roofline"]["node_visits_per_segment"]))
except Exception as e: print("   no json",e)
P
done
