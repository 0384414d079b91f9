#!/bin/bash
# ncu --set full of one origin-pass kernel on the bench workload, after the same command ran plain.
# usage: profile_kernel.sh [kernel regex, default swg_origin_block_kernel] [output name]
cd /root/repo; mkdir -p gpurun_out
K=${1:-swg_origin_block_kernel}; O=${2:-k2b_full}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/k2_plain.json 2> gpurun_out/k2_plain.err || { echo plain run failed; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/$O $CMD > gpurun_out/ncu_$O.log 2>&1
ls -la gpurun_out | tail -3
