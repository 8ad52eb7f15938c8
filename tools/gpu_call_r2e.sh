#!/bin/bash
# round 2: full GPU suite with the new end-to-end / output tests, then the bench line (device-resident + end-to-end at 512 streams)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nproc; free -g | head -2
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== bench.py"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "exit $?"; tail -c 1500 gpurun_out/r2e_bench.err; python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2e_bench.json").read().strip().splitlines()[-1])
    print(json.dumps({k: j[k] for k in ("value", "ms_per_step", "e2e", "stage_ms_per_step", "cpu_baseline")}, indent=1)[:3000])
    print("roofline", j["roofline"]["frac"], j["roofline"]["achieved"])
except Exception as e:
    print("no bench line", e)
PY
