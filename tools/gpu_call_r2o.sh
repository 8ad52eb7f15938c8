#!/bin/bash
# round 2: copy microbenchmark (what the memory system gives a pass-A-like copy) + ncu of both pass-A kernels on picture 12
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 300 tools/probe/copy_probe 2>&1 | tee gpurun_out/r2o_copy_probe.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"passA" -s 22 -c 2 -o gpurun_out/r2o_p12 python tools/prof_step.py 512 14 > gpurun_out/r2o_ncu12.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/r2o_ncu12.log
