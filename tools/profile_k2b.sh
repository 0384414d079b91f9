#!/bin/bash
# ncu --set full of the block-parallel origin pass (K2b) on the bench workload, after the same command ran plain.
cd /root/repo; mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/k2b_plain.json 2> gpurun_out/k2b_plain.err || { echo plain run failed; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:swg_origin_block_kernel -s 1 -c 1 -f -o gpurun_out/k2b_full $CMD > gpurun_out/ncu_k2b.log 2>&1
ls -la gpurun_out | tail -4
