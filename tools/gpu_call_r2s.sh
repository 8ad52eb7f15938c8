#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"passA" -s 22 -c 2 -o gpurun_out/r2s_p12 python tools/prof_step.py 512 14 > gpurun_out/r2s_ncu12.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/r2s_ncu12.log
