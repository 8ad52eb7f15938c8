#!/bin/bash
# GPU suite + stage timings of a 512-stream replay + per-picture stage times (one GPU call)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
echo "== parity"
timeout 1500 python -m pytest tests/test_gpu_synth.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
echo "== quick bench"
for i in 1 2; do
timeout 300 python tools/quick_bench.py 512 2 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['ms_per_pass_concurrent'],1), round(j['ms_per_pass'],1), {k: round(v,1) for k,v in j['stage_ms_per_pass'].items()}, j['watchdog'])"
done
echo "== per picture"
timeout 300 python tools/per_picture.py 512 > gpurun_out/per_picture.txt 2>&1; tail -1 gpurun_out/per_picture.txt
sed -n '2p;13p;41p;73p' gpurun_out/per_picture.txt
