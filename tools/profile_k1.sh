#!/bin/bash
# ncu evidence for K1 (run under gpurun, one GPU).  Outputs land in gpurun_out/ and are summarised into profiles/.
set -x
cd /root/repo
CMD="python bench.py --steps 1 --warmup 1 --genome-bp 240000 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo plain run failed; tail -20 gpurun_out/prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:swg_score_kernel -s 1 -c 1 -o gpurun_out/k1_full $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_list.log gpurun_out/ncu_full.log
ls -la gpurun_out
