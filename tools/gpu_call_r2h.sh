#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== parity"
timeout 1200 python -m pytest tests/test_gpu_synth.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for lib in libh264bsd_b200.so libvar_nofence.so; do
  echo "== quick bench 512, $lib"
  B200_LIB=$PWD/h264bsd_b200/$lib timeout 300 python tools/quick_bench.py 512 2 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['ms_per_pass_concurrent'], j['stage_ms_per_pass'], j['watchdog'])"
done
echo "== nofence parity"
B200_LIB=$PWD/h264bsd_b200/libvar_nofence.so timeout 900 python -m pytest tests/test_gpu_synth.py -m gpu -x -q -k "batched_engine or still" 2>&1 | tail -3
timeout 300 python tools/quick_bench.py 256 2 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(256, j['ms_per_pass_concurrent'], j['stage_ms_per_pass'])"
