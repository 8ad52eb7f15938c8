#!/bin/bash
# round 2: ncu of the two pass-A instances on a typical P picture (12) and on the all-copies picture (72)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"passAKernelT" -s 22 -c 2 -o gpurun_out/r2n_p12 python tools/prof_step.py 512 14 > gpurun_out/r2n_ncu12.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/r2n_ncu12.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"passAKernelT" -s 140 -c 2 -o gpurun_out/r2n_p72 python tools/prof_step.py 512 73 > gpurun_out/r2n_ncu72.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/r2n_ncu72.log
timeout 300 python tools/quick_bench.py 512 2 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['ms_per_pass_concurrent'],1), round(j['ms_per_pass'],1), {k: round(v,1) for k,v in j['stage_ms_per_pass'].items()}, j['watchdog'])"
