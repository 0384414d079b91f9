#!/bin/bash
# A/B of the library packing (run capacity in stripes) on the bench workload, after the parity tests.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py tests/test_gpu_fullsize.py -x -q --timeout 300 > gpurun_out/pack_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/pack_tests.log
tail -3 gpurun_out/pack_tests.log
for cap in 1 3 4 8 16; do
  echo "== RMGPU_PACK_RUN=$cap"
  RMGPU_PACK_RUN=$cap RMGPU_DEBUG_K1=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/pack_$cap.json 2> gpurun_out/pack_$cap.err
  grep "K1 dbg" gpurun_out/pack_$cap.err | tail -2
  python - <<PY
import json
b=json.loads(open("gpurun_out/pack_$cap.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f stage %s" % (b["value"], b["e2e"]["value"], {k: round(v,1) for k,v in b["stage_ms"].items()}))
PY
done
