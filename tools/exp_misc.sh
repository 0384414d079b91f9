#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py tests/test_gpu_fullsize.py -x -q --timeout 400 > gpurun_out/misc_tests.log 2>&1; tail -2 gpurun_out/misc_tests.log
for v in default w1 w4; do
  echo "== $v"
  L=""; [ $v != default ] && L=$PWD/repeatmasker_b200/_variants/librmgpu_$v.so
  RMGPU_LIB=$L RMGPU_DEBUG_K2=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/$v.json 2> gpurun_out/$v.err; grep "K2 dbg\] chain\|blocks: 992" gpurun_out/$v.err | tail -2
  python - $v <<'PY'
import json,sys
b=json.loads(open("gpurun_out/%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
print("value %.0f e2e %.0f stage %s cells %.3g" % (b["value"], b["e2e"]["value"], {k: round(v,1) for k,v in b["stage_ms"].items()}, b["origin_cells"]))
PY
done
