#!/bin/bash
# Round-2 evidence (run under gpurun, 1 GPU; all ncu runs of one call count as one profiler use):
#   0. the whole GPU test suite (with durations)
#   1. the bench command plain (its exit code gates everything else) and the --impl reference line
#   2. ncu launch list of the same command
#   3. ncu --set full of K1 on the 18-fragment bench launch itself, and of the other kernels of a step: the int32 score pass
#      of the wrapped triples (K2w), the block-parallel origin pass (both launches), the traceback (both teams), the row
#      kernel (K4) and the '*' / rank kernel
set -x
cd /root/repo
mkdir -p gpurun_out
if [ "$1" != "notests" ]; then
  timeout 1100 python -m pytest tests -x -q -m gpu --durations=12 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_gpu.log
  tail -18 gpurun_out/r2_pytest_gpu.log
fi
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_line.json 2> gpurun_out/r2_bench_line.err || { echo bench failed; tail -20 gpurun_out/r2_bench_line.err; exit 1; }
cat gpurun_out/r2_bench_line.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_line_reference.json 2> gpurun_out/r2_bench_line_reference.err; echo "reference arm rc=$?"
cat gpurun_out/r2_bench_line_reference.json
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-driver"
timeout 600 $CMD > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err || { echo plain run failed; tail -20 gpurun_out/r2_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"swg_score32_kernel|swg_origin_block_kernel|traceback_kernel|render_alignment_kernel|star_rank_kernel" -s 7 -c 7 -f -o gpurun_out/r2_k2k3k4_full $CMD > gpurun_out/r2_ncu_k2.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:swg_score_kernel -s 1 -c 1 -f -o gpurun_out/r2_k1_full $CMD > gpurun_out/r2_ncu_k1.log 2>&1
ls -la gpurun_out | tail -12
