#!/bin/bash
# Round-1 final evidence (run under gpurun, 1 GPU):
#  (1) plain bench (default command) -> gpurun_out/bench_final.json
#  (2) ncu launch list of the same command
#  (3) ncu --set full of K1 on a 240 kbp instance of the same workload (kernel replay of the 1 Mbp launch would have to
#      save/restore a >10 GB boundary ring ~40 times)
#  (4) dram bytes of one K1 launch at the full 1 Mbp size (2 metrics only)
set -x
cd /root/repo
CMD="python bench.py --steps 3 --warmup 3"
timeout 600 $CMD > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || { echo plain run failed; tail -20 gpurun_out/bench_final.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_list_final.log 2>&1
SMALL="python bench.py --steps 1 --warmup 1 --genome-bp 240000 --no-cpu-baseline"
timeout 200 $SMALL > /dev/null 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:swg_score_kernel -s 1 -c 1 -o gpurun_out/k1_full_final $SMALL > gpurun_out/ncu_full_final.log 2>&1
BIG="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:swg_score_kernel -s 1 -c 1 --csv --log-file gpurun_out/k1_dram_1mbp.csv $BIG > gpurun_out/ncu_dram.log 2>&1
ls -la gpurun_out | tail -8
