#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for k in 1 2 3; do timeout 300 python tools/_dbg.py > gpurun_out/dbg_split_$k.log 2>&1; echo "run $k rc=$?"; done
grep -v Traceback gpurun_out/dbg_split_1.log | tail -14
