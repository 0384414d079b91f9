cd /root/repo; mkdir -p gpurun_out
for v in 4 1; do
  RMGPU_K1_SEGMENTS=$v timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-driver 2>gpurun_out/job11_bench$v.err | tee gpurun_out/job11_bench$v.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('segments=$v', d['value'], d['ms_per_step'], d['stage_ms']['ms_score'], d['roofline']['frac'], d['hits'])"
done
tail -3 gpurun_out/job11_bench4.err
RMGPU_K1_SEGMENTS=3 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "score_pass or synthetic or long_subject or overflow or wrapped or kat or packing or alphabet" 2>&1 | tail -5
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 400 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu --durations=6 2>&1 | tail -12
