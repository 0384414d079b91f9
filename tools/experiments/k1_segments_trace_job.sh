cd /root/repo; mkdir -p gpurun_out
RMGPU_K1_TRACE=/root/repo/gpurun_out/k1_trace_s4.bin RMGPU_DEBUG_K1=1 RMGPU_K1_SEGMENTS=4 timeout 100 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-driver 2>gpurun_out/job13.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('segments=4', d['value'], d['ms_per_step'], d['stage_ms']['ms_score'], d['hits'])"
grep "K1 dbg" gpurun_out/job13.err | tail -2
ls -la gpurun_out/k1_trace_s4.bin
