#!/bin/bash
# round 2: full GPU suite, pass-A occupancy variants, stream groups side by side, bench line
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for lib in libh264bsd_b200.so libvar2_w4_b5.so libvar2_w8_b3.so; do
  echo "== quick bench 512, $lib"
  B200_LIB=$PWD/h264bsd_b200/$lib timeout 300 python tools/quick_bench.py 512 2 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['ms_per_pass_concurrent'], j['stage_ms_per_pass'])"
done
echo "== groups"
for cfg in "1 512 1" "1 512 2" "2 512 2" "2 512 4" "4 512 4"; do
  set -- $cfg
  echo "-- GRID_DIV=$1 groups=$3"
  B200_GRID_DIV=$1 timeout 300 python tools/group_bench.py $2 $3 2 0 2>&1 | tail -1
done
echo "== bench.py"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "exit $?"; tail -c 800 gpurun_out/r2f_bench.err; python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
    print(json.dumps({k: j[k] for k in ("value", "ms_per_step", "e2e", "stage_ms_per_step")}, indent=1)[:2500])
    print("roofline", j["roofline"]["frac"], j["roofline"]["achieved"])
except Exception as e:
    print("no bench line", e)
PY
