#!/bin/bash
# ncu --set full of the extend (both arms) and shade kernels in steady state of the bench command
mkdir -p gpurun_out
export PYTHONPATH=$PWD
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs"
timeout 600 $CMD > gpurun_out/plain_cand.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_wf_extend<' -s 150 -c 1 -f -o gpurun_out/r02_extend_cand $CMD > gpurun_out/ncu_ec.log 2>&1
echo "extend cand rc=$?"
timeout 600 $CMD --option 2=1 > gpurun_out/plain_exact.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_wf_extend<' -s 150 -c 1 -f -o gpurun_out/r02_extend_exact $CMD --option 2=1 > gpurun_out/ncu_ee.log 2>&1
echo "extend exact rc=$?"
timeout 600 $CMD > gpurun_out/plain_cand2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_wf_shade<' -s 150 -c 1 -f -o gpurun_out/r02_shade $CMD > gpurun_out/ncu_s.log 2>&1
echo "shade rc=$?"
tail -3 gpurun_out/ncu_ec.log
# destroy-time probe
python - <<'P'
import time, numpy as np
from surely_raytracing_b200 import BuiltScene, Scene
b = BuiltScene("c4")
out = np.zeros((800, 800, 3))
for k in range(6):
    t0 = time.perf_counter(); s = Scene(b); t1 = time.perf_counter(); s.render(0, 200, out=out); t2 = time.perf_counter(); s.close(); t3 = time.perf_counter()
    print("create %.2f render %.2f destroy %.2f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3))
P
ls -la gpurun_out/*.ncu-rep
