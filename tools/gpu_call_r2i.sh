#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== parity"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== per picture 512"
timeout 600 python tools/per_picture.py 512 > gpurun_out/r2i_per_picture.txt 2>&1
tail -1 gpurun_out/r2i_per_picture.txt
python - <<'PY'
import json
rows=[json.loads(l) for l in open("gpurun_out/r2i_per_picture.txt") if l.startswith('{"pic"')]
for r in rows[:4]+rows[40:43]: print(r)
import statistics
p=[r for r in rows if r["nB"]<8000]
for k in ("recon","deblock","recon_intra","strength","border"): print(k, "P-picture mean", round(statistics.mean(r[k] for r in p),3), "IDR", [round(r[k],2) for r in rows if r["nB"]>=8000])
PY
