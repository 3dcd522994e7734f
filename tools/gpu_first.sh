#!/bin/bash
# GPU pass: smoke, GPU parity tests, a short bench, then ncu (launch list + one full capture)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)|Socket" >> gpurun_out/nproc.txt
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
echo "== ncu"
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_render_mega -s 1 -c 1 -o gpurun_out/prof_mega python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out
