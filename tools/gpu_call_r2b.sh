#!/bin/bash
# round 2: the strip-layout engine on hardware for the first time -- parity, stage times, one ncu capture of pass A
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q -rfEX > gpurun_out/r2b_gpu_tests.txt 2>&1
echo "exit $?"; tail -n 25 gpurun_out/r2b_gpu_tests.txt
echo "== quick bench 256 / 512"
timeout 300 python tools/quick_bench.py 256 2 2>&1 | tail -2
timeout 300 python tools/quick_bench.py 512 2 2>&1 | tail -2
echo "== ncu pass A"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:passAKernel -s 4 -c 1 -o gpurun_out/r2b_passA python tools/prof_step.py 256 6 > gpurun_out/r2b_ncu.log 2>&1
echo "exit $?"; tail -n 3 gpurun_out/r2b_ncu.log
