#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
grep -m1 "model name" /proc/cpuinfo; grep -o -w -E "avx2|bmi2|avx512f" /proc/cpuinfo | sort | uniq -c | head -3
echo "== parser: base vs x86-64-v3 host build"
for i in 1 2; do for l in libvar_base.so libh264bsd_b200.so; do echo -n "$l: "; B200_LIB=$PWD/h264bsd_b200/$l python tools/parse_scale.py 16 2>&1 | tail -1; done; done
echo "== parity"
timeout 1200 python -m pytest tests/test_gpu_synth.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
echo "== quick bench"
timeout 300 python tools/quick_bench.py 512 2 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['ms_per_pass_concurrent'], j['stage_ms_per_pass'], j['watchdog'])"
