#!/bin/bash
# Round-2 multi-GPU evidence on one 8-GPU box (gpurun --gpus 8).  Everything lands in gpurun_out/r2_scale_*.
#   weak    the driver's own scaling command: BASELINE configs[1], one genome of N x 1 Mbp, fragment b -> rank b mod N,
#           library replicated, hit records gathered on rank 0 inside the timed region
#   strong  BASELINE configs[2] (1,500 families, -s parameters) on a FIXED 8 Mbp slice of its genome (144 fragments)
#   product the unmodified driver, RepeatMasker -engine gpu -pa 32, through rmgpu_server over all 8 GPUs:
#           configuration 3 at full size (46 Mbp) and a 2 Mbp slice of configuration 5 (20,000 families, 24 Mbp library)
cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > gpurun_out/r2_scale_gpus.txt
for N in 1 2 4 8; do
  if [ $N = 1 ]; then
    python bench.py --gpus 1 --steps 3 --warmup 2 --no-cpu-baseline --no-driver > gpurun_out/r2_scale_weak_1.json 2> gpurun_out/r2_scale_weak_1.err
  else
    $TR --nproc-per-node $N --master-port $((29500 + N)) bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r2_scale_weak_$N.json 2> gpurun_out/r2_scale_weak_$N.err
  fi
  tail -c 300 gpurun_out/r2_scale_weak_$N.json | cut -c1-10 > /dev/null
done
for N in 1 2 4 8; do
  ARGS="--gpus $N --steps 2 --warmup 1 --workload config3 --scaling strong --genome-bp 8000000 --no-cpu-baseline --no-driver"
  if [ $N = 1 ]; then
    python bench.py $ARGS > gpurun_out/r2_scale_strong_1.json 2> gpurun_out/r2_scale_strong_1.err
  else
    $TR --nproc-per-node $N --master-port $((29600 + N)) bench.py $ARGS > gpurun_out/r2_scale_strong_$N.json 2> gpurun_out/r2_scale_strong_$N.err
  fi
done
python tools/run_driver_config2.py --families 1500 --seed 1002 --mbp 46 --extra=-s --pa 32 --server-gpus 0,1,2,3,4,5,6,7 > gpurun_out/r2_scale_driver_config3_8gpu.txt 2>&1
python tools/run_driver_config2.py --families 1500 --seed 1002 --mbp 46 --extra=-s --pa 32 --server-gpus 0 > gpurun_out/r2_scale_driver_config3_1gpu.txt 2>&1
python tools/run_driver_config2.py --families 20000 --seed 1005 --mbp 2 --pa 32 --server-gpus 0,1,2,3,4,5,6,7 > gpurun_out/r2_scale_driver_config5_slice_8gpu.txt 2>&1
grep -h '"value"' gpurun_out/r2_scale_weak_*.json gpurun_out/r2_scale_strong_*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['scaling'], d['n_gpus'], round(d['value']), 'GCUPS', round(d['ms_per_step'], 1), 'ms/step', d['hits'], 'hits', 'e2e', round(d['e2e']['value']))
"
tail -4 gpurun_out/r2_scale_driver_*.txt
