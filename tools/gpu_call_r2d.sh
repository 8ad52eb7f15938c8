#!/bin/bash
# round 2: the row-walking filter on hardware
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== quick bench 512 / 256"
timeout 300 python tools/quick_bench.py 512 2 2>&1 | tail -1
timeout 300 python tools/quick_bench.py 256 2 2>&1 | tail -1
echo "== ncu filter + pass A"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"passAKernel|deblockKernel|reconIntraKernel" -s 12 -c 3 -o gpurun_out/r2d_prof python tools/prof_step.py 256 6 > gpurun_out/r2d_ncu.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/r2d_ncu.log
