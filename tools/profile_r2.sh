#!/bin/bash
# Round-2 evidence (run under gpurun, 1 GPU; all ncu runs of one call count as one profiler use):
#   1. the bench command plain (its exit code gates everything else)
#   2. ncu launch list of the same command
#   3. ncu --set full of K1 on the 18-fragment bench launch itself, of the block-parallel origin pass, of the traceback
#      (block teams) and of the row kernel (K4)
set -x
cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-driver"
timeout 600 $CMD > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err || { echo plain run failed; tail -20 gpurun_out/r2_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:swg_score_kernel -s 1 -c 1 -f -o gpurun_out/r2_k1_full $CMD > gpurun_out/r2_ncu_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"swg_origin_block_kernel|traceback_kernel|render_alignment_kernel|star_rank_kernel" -s 4 -c 5 -f -o gpurun_out/r2_k2k3k4_full $CMD > gpurun_out/r2_ncu_k2.log 2>&1
ls -la gpurun_out | tail -8
