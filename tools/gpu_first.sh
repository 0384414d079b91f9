set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
cd /root/repo
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 120 python -c "
from repeatmasker_b200 import engine
c,i=engine.measure_dpx_peak(0); print('DPX peak GCUPS', c/1e9, 'VIADDMNMX warp-instr/s', i)
"
