#!/bin/bash
# The profiles of the benchmarked build (one GPU call): launch list of the bench command, ncu --set full of every kernel at 512
# streams (a P picture and the IDR picture), the output kernels.  Reports go to gpurun_out/$TAG_*; tools/ncu_summary.py turns them
# into the text files under profiles/.
#   tools/gpurun_built.sh --timeout 3600 -- 'bash tools/gpu_calls/profile.sh r02'
cd "$(dirname "$0")/../.." || exit 1
TAG="${1:-prof}"
mkdir -p gpurun_out
python -c "import bench; print(bench.kernel_source_sha16())" > gpurun_out/${TAG}_sha.txt; cat gpurun_out/${TAG}_sha.txt
echo "== launch list of the bench command (warm-up passes: the same launches as the timed ones)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 620 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --cpu-seconds 1 > gpurun_out/${TAG}_launches.log 2>&1
echo "exit $?"; wc -l gpurun_out/${TAG}_launches.csv
echo "== ncu --set full, 512 streams, one P picture (picture 12)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"passAKernel|passAMultiKernel|deblockKernel|reconIntraKernel|strengthKernel|borderKernel" -s 70 -c 6 -o gpurun_out/${TAG}_prof512 python tools/prof_step.py 512 14 > gpurun_out/${TAG}_ncu.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/${TAG}_ncu.log
echo "== ncu --set full, 512 streams, the IDR picture (intra pass, filter)"
timeout 1200 ncu --set full --clock-control none -k regex:"deblockKernel|reconIntraKernel" -c 2 -o gpurun_out/${TAG}_prof512_idr python tools/prof_step.py 512 1 > gpurun_out/${TAG}_ncu_idr.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/${TAG}_ncu_idr.log
echo "== convert + pack kernels"
timeout 600 ncu --set full --clock-control none -k regex:"convertFrameKernel|packKernel" -c 2 -o gpurun_out/${TAG}_prof_out python tools/convert_bench.py > gpurun_out/${TAG}_ncu_out.log 2>&1
echo "exit $?"; tail -n 3 gpurun_out/${TAG}_ncu_out.log
