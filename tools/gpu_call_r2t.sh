#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
qb() { timeout 300 python tools/quick_bench.py 512 2 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['ms_per_pass_concurrent'],1), round(j['ms_per_pass'],1), {k: round(v,1) for k,v in j['stage_ms_per_pass'].items()}, j['watchdog'])"; }
for v in 4_17 4_23; do echo "== CTAs_chunk $v"; B200_LIB=$PWD/h264bsd_b200/libexp_$v.so qb; done
