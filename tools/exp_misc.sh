#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py tests/test_gpu_fullsize.py -x -q --timeout 400 > gpurun_out/misc_tests.log 2>&1; tail -2 gpurun_out/misc_tests.log
RMGPU_DEBUG_K1=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/k1.json 2> gpurun_out/k1.err; grep "K1 dbg" gpurun_out/k1.err | tail -2
python - <<'PY'
import json
b=json.loads(open("gpurun_out/k1.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f stage %s" % (b["value"], b["e2e"]["value"], {k: round(v,1) for k,v in b["stage_ms"].items()}))
PY
