#!/bin/bash
# stream groups side by side (tools/group_bench.py) against the single batch, in ONE gpurun call -> gpurun_out/ab_groups.txt
#   gpurun --timeout 900 -- 'bash tools/ab_groups.sh'
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
out=gpurun_out/ab_groups.txt
: > "$out"
run() {  # env-assignments... -- args of group_bench.py
    echo "== $*" >> "$out"
    local envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    env "${envs[@]}" timeout 240 python tools/group_bench.py "$@" >> "$out" 2>&1 || echo "FAILED ($?)" >> "$out"
}
run B200_GRID_DIV=1 -- 512 1 2 0          # today's schedule, measured the same way
run B200_GRID_DIV=1 -- 512 2 2 0
run B200_GRID_DIV=2 -- 512 2 2 0
run B200_GRID_DIV=1 -- 512 4 2 0
run B200_GRID_DIV=2 -- 512 4 2 0
run B200_GRID_DIV=4 -- 512 4 2 0
run B200_GRID_DIV=2 -- 512 4 2 1          # groups one picture apart
run B200_GRID_DIV=4 -- 512 8 2 0
cat "$out"
