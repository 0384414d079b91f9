#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -x -q --timeout 300 > gpurun_out/misc_tests.log 2>&1; tail -2 gpurun_out/misc_tests.log
RMGPU_DEBUG_HOST=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/host.json 2> gpurun_out/host.err
grep "set_library\|set_queries" gpurun_out/host.err | tail -2
python - <<'PY'
import json
b=json.loads(open("gpurun_out/host.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f (%.1f vs %.1f ms) stage %s" % (b["value"], b["e2e"]["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], {k: round(v,1) for k,v in b["stage_ms"].items()}))
PY
