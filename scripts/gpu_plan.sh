#!/bin/bash
# usage (under gpurun): scripts/gpu_plan.sh "<size>" ...   -- pass time of a channel of that size for several per-item overheads of the plan's cost model
for sz in "$@"; do
for ov in 0 50 100 200 400; do
  export SB_PLAN_OVERHEAD=$ov
  python bench.py --workload c5s --size $sz --steps 4 --warmup 2 --no-cpu --no-verify --e2e-depth 1 > gpurun_out/plan_$ov.json 2> gpurun_out/plan_$ov.err
  python -c "
import json; d=json.loads(open('gpurun_out/plan_$ov.json').read().strip().splitlines()[-1]); print('$sz ov $ov', round(d['value'],1), 'pass', round(d['roofline']['avg_launch_ms'],4), d['config']['rb_plan'])"
done; done
