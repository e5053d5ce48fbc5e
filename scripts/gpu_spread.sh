#!/bin/bash
# usage (under gpurun --gpus 2): scripts/gpu_spread.sh -- spread of the pass durations, 1 GPU vs 2 slabs
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/spread_n1.json 2> gpurun_out/spread_n1.err
SB_FIN_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/spread_n2.json 2> gpurun_out/spread_n2.err
grep "finalize trace" gpurun_out/spread_n2.err
python - <<PY
import json
for n in (1,2):
    d=json.loads(open(f"gpurun_out/spread_n{n}.json").read().strip().splitlines()[-1])
    print(f"N={n} Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "sor", round(d["sor"]["ms_per_tick"],3), "pass ms", round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["launch_ms_min_median_max"], d["config"]["rb_plan"])
PY
