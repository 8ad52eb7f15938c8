#!/bin/bash
# round 2: timing experiments (not bit-exact builds): pass A without luma / chroma / residual / copies / inter; filter wait back-off
cd "$(dirname "$0")/.." || exit 1
qb() { timeout 300 python tools/quick_bench.py 512 2 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['ms_per_pass_concurrent'],1), round(j['ms_per_pass'],1), {k: round(v,1) for k,v in j['stage_ms_per_pass'].items()}, j['watchdog'])"; }
echo "== default"; qb
for e in 1 2 3 4 5; do echo "== EXP $e (1 no luma interpolation, 2 no chroma, 3 no residual, 4 no copies, 5 no inter)"; B200_LIB=$PWD/h264bsd_b200/libexp$e.so qb; done
i=0; for cfg in "100 1000 0" "250 2000 0" "100 1000 4" "100 1000 8" "20 640 6"; do i=$((i+1)); echo "== filter wait: ns0 max slack = $cfg"; B200_LIB=$PWD/h264bsd_b200/libdb$i.so qb; done
