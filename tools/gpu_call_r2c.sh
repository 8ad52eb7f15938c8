#!/bin/bash
# round 2: pass A with out-of-line helpers (code size), filter with two macroblocks per warp; register-budget variant
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== parity (synthetic + fixtures, quick subset)"
timeout 900 python -m pytest tests/test_gpu_synth.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for lib in libh264bsd_b200.so libvar_mb2.so; do
  echo "== quick bench 512, $lib"
  B200_LIB=$PWD/h264bsd_b200/$lib timeout 300 python tools/quick_bench.py 512 2 2>&1 | tail -1
done
echo "== ncu pass A + filter"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"passAKernel|deblockKernel" -s 8 -c 2 -o gpurun_out/r2c_prof python tools/prof_step.py 256 6 > gpurun_out/r2c_ncu.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/r2c_ncu.log
