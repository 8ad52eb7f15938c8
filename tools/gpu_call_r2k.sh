#!/bin/bash
# round 2: profiles of the benchmarked build -- launch list of the bench command, ncu --set full of every kernel at 512 streams
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
python -c "import bench; print(bench.kernel_source_sha16())" > gpurun_out/r2k_sha.txt; cat gpurun_out/r2k_sha.txt
echo "== launch list of the bench command (warm-up passes: the same launches as the timed ones)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 520 --csv --log-file gpurun_out/r2k_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --cpu-seconds 1 > gpurun_out/r2k_launches.log 2>&1
echo "exit $?"; wc -l gpurun_out/r2k_launches.csv
echo "== ncu --set full, 512 streams, one P picture"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"passAKernelT|deblockKernel|reconIntraKernel|strengthKernel|borderKernel" -s 30 -c 6 -o gpurun_out/r2k_prof512 python tools/prof_step.py 512 8 > gpurun_out/r2k_ncu.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/r2k_ncu.log
echo "== ncu --set full, 512 streams, the IDR picture (intra pass, filter)"
timeout 1200 ncu --set full --clock-control none -k regex:"deblockKernel|reconIntraKernel" -c 2 -o gpurun_out/r2k_prof512_idr python tools/prof_step.py 512 1 > gpurun_out/r2k_ncu_idr.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/r2k_ncu_idr.log
echo "== convert + pack kernels"
timeout 600 ncu --set full --clock-control none -k regex:"convertFrameKernel|packKernel" -c 2 -o gpurun_out/r2k_prof_out python tools/convert_bench.py > gpurun_out/r2k_ncu_out.log 2>&1
echo "exit $?"; tail -n 3 gpurun_out/r2k_ncu_out.log
