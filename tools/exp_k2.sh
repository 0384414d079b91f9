#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py tests/test_gpu_fullsize.py -x -q --timeout 300 > gpurun_out/k2_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/k2_tests.log
tail -4 gpurun_out/k2_tests.log
RMGPU_DEBUG_K2=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/k2dbg.json 2> gpurun_out/k2dbg.err
grep "dbg\]" gpurun_out/k2dbg.err | tail -4
python - <<PY
import json
b=json.loads(open("gpurun_out/k2dbg.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f stage %s redo %s cells %.3g" % (b["value"], b["e2e"]["value"], {k: round(v,1) for k,v in b["stage_ms"].items()}, b["origin_redo"], b["origin_cells"]))
PY
