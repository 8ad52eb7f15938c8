#!/bin/bash
# round 2, first call: what round 1 left unmeasured (parser scaling on the box, copy-pass variants, stream groups)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi -L; nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; lscpu | grep -i -E 'numa|model name|socket' ; nvidia-smi topo -m 2>/dev/null | head -20
echo "== parse_scale"
timeout 300 python tools/parse_scale.py 1 8 16 > gpurun_out/r2a_parse_scale.txt 2>&1
cat gpurun_out/r2a_parse_scale.txt
echo "== copy-pass variants"
bash tools/ab_copy.sh
echo "== stream groups side by side"
bash tools/ab_groups.sh
