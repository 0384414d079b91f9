#!/bin/bash
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.0f e2e %.0f K1 %.1f ms stage %s" % (b["value"], b["e2e"]["value"], b["stage_ms"]["ms_score"], {k: round(v,1) for k,v in b["stage_ms"].items()}))
PY
}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -x -q --timeout 300 > gpurun_out/misc_tests.log 2>&1; tail -2 gpurun_out/misc_tests.log
timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/vnew.json 2> gpurun_out/vnew.err
summ gpurun_out/vnew.json
