#!/bin/bash
mkdir -p gpurun_out
RMGPU_K2_SERIAL=1 RMGPU_DEBUG_K2=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/k4.json 2> gpurun_out/k4.err; grep "K2 dbg" gpurun_out/k4.err | tail -3
echo "== overlapped"
RMGPU_DEBUG_K2=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/k4.json 2> gpurun_out/k4.err; grep "K2 dbg" gpurun_out/k4.err | tail -4
python - <<'PY'
import json
b=json.loads(open("gpurun_out/k4.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f stage %s" % (b["value"], b["e2e"]["value"], {k: round(v,1) for k,v in b["stage_ms"].items()}))
PY
