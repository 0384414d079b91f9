#!/bin/bash
# One gpurun call: the whole GPU suite (with durations), then the plain bench line, then the ncu launch list of the same command.
cd /root/repo; mkdir -p gpurun_out
timeout 1100 python -m pytest tests -x -q -m gpu --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
CMD="python bench.py --steps 3 --warmup 3"
timeout 600 $CMD > gpurun_out/bench_line.json 2> gpurun_out/bench_line.err || { echo plain run failed; tail -20 gpurun_out/bench_line.err; exit 1; }
cat gpurun_out/bench_line.json
CMD2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-driver"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD2 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
