cd /root/repo; mkdir -p gpurun_out
# K1 cell 4 (default) against cell 5 on the bench, then the K1-relevant parity tests with the default build
for v in 4 5; do
  RMGPU_K1_CELL=$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-driver 2>gpurun_out/job8_bench$v.err | tee gpurun_out/job8_bench$v.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cell=$v', d['value'], d['ms_per_step'], d['stage_ms']['ms_score'], d['roofline']['frac'], d['hits'])"
done
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --durations=5 -k "score_pass or synthetic or long_subject or overflow or wrapped or matrix_sweep or kat or packing or alphabet" 2>&1 | tail -12
