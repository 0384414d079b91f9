#!/bin/bash
set -x
cd /root/repo
CMD="python bench.py --steps 1 --warmup 1 --genome-bp 240000 --no-cpu-baseline"
timeout 200 $CMD > gpurun_out/prof_plain.log 2>&1 || { echo plain run failed; tail -20 gpurun_out/prof_plain.log; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:swg_origin_block_kernel -s 1 -c 1 -o gpurun_out/k2b_full $CMD > gpurun_out/ncu_full_k2b2.log 2>&1
ls -la gpurun_out | grep k2b
