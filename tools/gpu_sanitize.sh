#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize.log 2>&1
echo "rc=$?"; grep -v "^$" gpurun_out/sanitize.log | head -60
