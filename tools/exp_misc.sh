#!/bin/bash
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.0f e2e %.0f stage %s" % (b["value"], b["e2e"]["value"], {k: round(v,1) for k,v in b["stage_ms"].items()}))
PY
}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 > gpurun_out/misc_tests.log 2>&1; tail -2 gpurun_out/misc_tests.log
RMGPU_DEBUG_HOST=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/host.json 2> gpurun_out/host.err
grep "host dbg" gpurun_out/host.err | tail -12; summ gpurun_out/host.json
for u in 1 4; do
  echo "== K1 unroll $u"
  RMGPU_LIB=$PWD/repeatmasker_b200/_variants/librmgpu_u$u.so timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/u$u.json 2> gpurun_out/u$u.err
  summ gpurun_out/u$u.json
done
