#!/bin/bash
# round 2: do the copy kernel and the computing kernel of pass A share the SMs when pictures are launched back to back?
cd "$(dirname "$0")/.." || exit 1
qb() { timeout 300 python tools/quick_bench.py 512 2 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['ms_per_pass_concurrent'],1), round(j['ms_per_pass'],1), {k: round(v,1) for k,v in j['stage_ms_per_pass'].items()}, j['watchdog'])"; }
echo "== default (5 CTAs/SM of 96 registers + 1 copy CTA)"; qb
echo "== B200_PASSA_CTAS=4"; B200_PASSA_CTAS=4 qb
echo "== B200_PASSA_CTAS=4 B200_COPY_CTAS=2"; B200_PASSA_CTAS=4 B200_COPY_CTAS=2 qb
echo "== 88 registers"; B200_LIB=$PWD/h264bsd_b200/libreg88.so qb
echo "== 80 registers"; B200_LIB=$PWD/h264bsd_b200/libreg80.so qb
echo "== 80 registers, B200_COPY_CTAS=2"; B200_COPY_CTAS=2 B200_LIB=$PWD/h264bsd_b200/libreg80.so qb
