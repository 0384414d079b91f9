cd /root/repo; mkdir -p gpurun_out
for w in 4 1000; do
  RMGPU_DEBUG_K1=1 RMGPU_K1_SEGWIN=$w timeout 100 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-driver 2>gpurun_out/job14_$w.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('window rounds=$w', d['value'], d['ms_per_step'], d['stage_ms']['ms_score'], d['roofline']['frac'], d['hits'])"
  grep "K1 dbg" gpurun_out/job14_$w.err | tail -2
done
RMGPU_K1_SEGMENTS=8 RMGPU_K1_SEGWIN=1000 timeout 100 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-driver 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('8 segments, one window', d['value'], d['ms_per_step'], d['stage_ms']['ms_score'], d['roofline']['frac'], d['hits'])"
