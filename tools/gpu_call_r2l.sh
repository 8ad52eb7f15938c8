#!/bin/bash
# round 2: do kernels of different streams share SMs once their shared-memory carveouts agree?  pass-A instances side by side; stream groups
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
qb() { timeout 300 python tools/quick_bench.py 512 2 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['ms_per_pass_concurrent'],1), round(j['ms_per_pass'],1), {k: round(v,1) for k,v in j['stage_ms_per_pass'].items()}, j['watchdog'])"; }
echo "== default"; qb
echo "== CARVEOUT=100"; B200_CARVEOUT=100 qb
echo "== CARVEOUT=50"; B200_CARVEOUT=50 qb
for sp in "3 2" "4 1" "4 2" "5 5"; do
  echo "== SPLIT_A=$sp"; B200_SPLIT_A="$sp" qb
  echo "== SPLIT_A=$sp CARVEOUT=100"; B200_SPLIT_A="$sp" B200_CARVEOUT=100 qb
done
echo "== groups, CARVEOUT=100"
for cfg in "1 512 1" "1 512 2" "2 512 2" "2 512 4"; do
  set -- $cfg
  echo "-- GRID_DIV=$1 groups=$3"
  B200_CARVEOUT=100 B200_GRID_DIV=$1 timeout 300 python tools/group_bench.py $2 $3 2 0 2>&1 | tail -1
done
echo "-- GRID_DIV=2 groups=2 stagger 1"
B200_CARVEOUT=100 B200_GRID_DIV=2 timeout 300 python tools/group_bench.py 512 2 2 1 2>&1 | tail -1
