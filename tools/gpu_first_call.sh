#!/bin/bash
# Everything round 1 left unmeasured, in ONE gpurun call (about 25 minutes of box time):
#   gpurun --timeout 2400 -- 'bash tools/gpu_first_call.sh'
# 1. the GPU suite with the xfail-marked tests reported as what they really do (concealKernel, the copy-pass variants)
# 2. bench.py (end-to-end figure with the faster host parser; device figure must not have moved: default kernels' SASS is unchanged)
# 3. host parser scaling on the box's host cores
# 4. copy-pass variants: parity + stage timings (tools/ab_copy.sh)
# 5. launch list of the bench command (kernel shares of a step) for profiles/
# Every step has its own timeout and writes to gpurun_out/; a failing step does not stop the next one.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== 1. pytest -m gpu --runxfail"
timeout 900 python -m pytest tests -m gpu -q -rfEX --runxfail > gpurun_out/first_gpu_tests.txt 2>&1
echo "exit $?"; tail -n 15 gpurun_out/first_gpu_tests.txt
echo "== 2. bench.py"
timeout 600 python bench.py > gpurun_out/first_bench.json 2> gpurun_out/first_bench.err
echo "exit $?"; tail -c 3000 gpurun_out/first_bench.json
echo "== 3. parse_scale"
timeout 300 python tools/parse_scale.py 1 8 16 > gpurun_out/first_parse_scale.txt 2>&1
cat gpurun_out/first_parse_scale.txt
echo "== 4. copy-pass variants"
bash tools/ab_copy.sh
echo "== 4b. stream groups side by side"
bash tools/ab_groups.sh
echo "== 5. launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/first_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --cpu-seconds 1 > gpurun_out/first_launches.log 2>&1
echo "exit $?"
