#!/bin/bash
# round 2: ncu captures of the current kernels (source counters), quick bench
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 300 python tools/quick_bench.py 512 2 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"passAKernelT|deblockKernel|reconIntraKernel|strengthKernel" -s 15 -c 5 -o gpurun_out/r2g_prof python tools/prof_step.py 256 6 > gpurun_out/r2g_ncu.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/r2g_ncu.log
