#!/bin/bash
# second ncu capture of round 1 (after K1/K2 restructuring): launch list + full sets of K1 and K2
set -x
cd /root/repo
CMD="python bench.py --steps 1 --warmup 1 --genome-bp 240000 --no-cpu-baseline"
timeout 200 $CMD > gpurun_out/prof_plain.log 2>&1 || { echo plain run failed; tail -20 gpurun_out/prof_plain.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_b.csv $CMD > gpurun_out/ncu_list_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:swg_score_kernel -s 1 -c 1 -o gpurun_out/k1_full_b $CMD > gpurun_out/ncu_full_k1b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:swg_origin_kernel -s 1 -c 1 -o gpurun_out/k2_full_b $CMD > gpurun_out/ncu_full_k2b.log 2>&1
ls -la gpurun_out | tail -8
