cd /root/repo; mkdir -p gpurun_out
# K1 cell 3 (chain-2, four alu-pipe instructions) against the shipped cell on the bench, then the K1-relevant parity tests with it
for v in 3 5; do
  RMGPU_K1_CELL=$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-driver 2>gpurun_out/job9_bench$v.err | tee gpurun_out/job9_bench$v.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cell=$v', d['value'], d['ms_per_step'], d['stage_ms']['ms_score'], d['roofline']['frac'], d['hits'])"
done
RMGPU_K1_CELL=3 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "score_pass or synthetic or long_subject or overflow or wrapped or matrix_sweep or kat" 2>&1 | tail -4
