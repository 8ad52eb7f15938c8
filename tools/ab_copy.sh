#!/bin/bash
# A/B of the copy-pass variants in ONE gpurun call (round 2, first thing): parity of every variant in a process of its own, then
# the per-stage times of a 256-stream 1080p replay with each knob setting -> gpurun_out/ab_copy.txt.
#   gpurun --timeout 900 -- 'bash tools/ab_copy.sh'
# Read stage_ms_per_pass.recon_copy (serialised per-kernel pass) and ms_per_pass_concurrent (the step as the bench times it).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
out=gpurun_out/ab_copy.txt
: > "$out"
echo "== parity (tests/test_gpu_synth.py -k copy_variants, --runxfail)" >> "$out"
timeout 600 python -m pytest tests/test_gpu_synth.py -q -k copy_variants --runxfail -x >> "$out" 2>&1
echo "exit code $?" >> "$out"
for knobs in "B200_COPY_VARIANT=0" "B200_COPY_VARIANT=1" "B200_COPY_VARIANT=2" \
             "B200_COPY_VARIANT=2 B200_COPY_RUNS=8" \
             "B200_COPY_BULK=1" "B200_COPY_BULK=1 B200_COPY_BULK_RUNS=8" "B200_COPY_BULK=1 B200_COPY_BULK_RUNS=4"; do
    echo "== $knobs" >> "$out"
    # shellcheck disable=SC2086
    env $knobs timeout 300 python tools/quick_bench.py 256 2 >> "$out" 2>&1 || echo "FAILED ($?)" >> "$out"
done
tail -n 40 "$out"
