#!/bin/bash
# Round-1 final evidence (run under gpurun, 1 GPU), one profiler pass per call:
#   bench  plain bench (default command) -> gpurun_out/bench_final.json, then the ncu launch list of the same command
#   full   ncu --set full of K1 on a 240 kbp instance of the same workload (kernel replay of the 1 Mbp launch would have
#          to save/restore a ~10 GB boundary ring ~40 times), after the same command ran plain
#   dram   dram bytes of one K1 launch at the full 1 Mbp size (2 metrics only)
set -x
cd /root/repo
mkdir -p gpurun_out
case "$1" in
bench)
  CMD="python bench.py --steps 3 --warmup 3"
  timeout 600 $CMD > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || { echo plain run failed; tail -20 gpurun_out/bench_final.err; exit 1; }
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_list_final.log 2>&1
  ;;
full)
  SMALL="python bench.py --steps 1 --warmup 1 --genome-bp 240000 --no-cpu-baseline"
  timeout 200 $SMALL > gpurun_out/small_plain.json 2> gpurun_out/small_plain.err || { echo plain run failed; exit 1; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:swg_score_kernel -s 1 -c 1 -f -o gpurun_out/k1_full_final $SMALL > gpurun_out/ncu_full_final.log 2>&1
  ;;
dram)
  BIG="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
  timeout 300 $BIG > gpurun_out/big_plain.json 2> gpurun_out/big_plain.err || { echo plain run failed; exit 1; }
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:swg_score_kernel -s 1 -c 1 --csv --log-file gpurun_out/k1_dram_1mbp.csv $BIG > gpurun_out/ncu_dram.log 2>&1
  ;;
*) echo "usage: $0 bench|full|dram"; exit 2 ;;
esac
ls -la gpurun_out | tail -6
