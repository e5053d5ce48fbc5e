#!/bin/bash
# usage (under gpurun --gpus 8): scripts/gpu_scale.sh <tag> "<workload> <N list>" ...
# e.g. scripts/gpu_scale.sh r2 "c5s 1 2 4 8" "c3 1 2 4" "c4 1 2 4 8"
# one bench line per (workload, N) into gpurun_out/<tag>_scale_<workload>_n<N>.json
TAG=$1; shift
PORT=29600
for spec in "$@"; do
  set -- $spec; W=$1; shift
  for N in "$@"; do
    OUT=gpurun_out/${TAG}_scale_${W}_n${N}
    PORT=$((PORT+1))
    if [ "$N" = 1 ]; then
      timeout 600 python bench.py --gpus 1 --workload $W --steps 5 --warmup 3 --no-cpu --no-verify > $OUT.json 2> $OUT.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus $N --workload $W --steps 5 --warmup 3 --no-cpu --no-verify > $OUT.json 2> $OUT.err
    fi
    python - <<PY
import json
try:
    d=json.loads(open("$OUT.json").read().strip().splitlines()[-1])
    r=d["roofline"] or {}
    print("$W N=$N", d["config"]["grid"], d["scaling"], "Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "sor ms", round(d["sor"]["ms_per_tick"],3), "pass ms", round(r.get("avg_launch_ms",0),4), "plan", d["config"]["rb_plan"], "e2e", round(d["e2e"]["value"],1))
except Exception as e:
    print("$W N=$N failed", e); print(open("$OUT.err").read()[-800:])
PY
  done
done
