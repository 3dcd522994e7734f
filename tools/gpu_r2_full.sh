#!/bin/bash
# full single-GPU pass: smoke, every gpu test, the bench line (with cpu baseline + configs block), the reference arm, the c5 curve
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -q --durations=6 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|pixel-centre ids|variant [01]:" gpurun_out/pytest_gpu.log | tail -20
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; grep -E "step_ms|e2e \(" gpurun_out/bench.err | cut -c1-220
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.1f e2e %.1f ms/step %.2f frac %.3f useful %.3f image_check %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["frac"],r["useful_instruction_fraction"],json.dumps(d["image_check"])[:400]))
print("cpu_baseline", d.get("cpu_baseline"))
for c in d.get("configs",[]): print("  ", c["workload"], "value %.1f e2e %.1f ms %.2f launches %d frac %.3f check %s"%(c["value"],c["e2e"],c["ms"],c["gpu_launches"],c["roofline_frac"],(c["image_check"] or {}).get("accepted")))
for k,v in r["kernels"].items(): print("  ", k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
P
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "reference arm rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 900 python bench.py --workload c5 --curve > gpurun_out/curve_c5.json 2> gpurun_out/curve_c5.err; echo "curve rc=$?"; tail -3 gpurun_out/curve_c5.err | cut -c1-200
