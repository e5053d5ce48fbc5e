#!/bin/bash
# usage (under gpurun --gpus N): scripts/gpu_scale.sh "<N list>" [bench args]  -- the driver's scaling run
NS=$1; shift
mkdir -p gpurun_out
for N in $NS; do
  if [ "$N" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "N=$N rc=$?"; tail -1 gpurun_out/scale_n$N.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('  Mcs/s',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'grid',d['config']['grid'],'plan',d['config'].get('rb_plan'),'pass ms',round(d['roofline']['avg_launch_ms'],4),'e2e',round(d['e2e']['value'],1))
except Exception as e: print('  failed',e)
"; tail -3 gpurun_out/scale_n$N.err
done
