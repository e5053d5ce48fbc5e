#!/bin/bash
# usage (under gpurun --gpus 2): scripts/gpu_ww.sh "<weights>" -- 2-slab bench per wall weight
mkdir -p gpurun_out
for w in $1; do
SB_WALL_WEIGHT=$w SB_FIN_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/ww_$w.json 2> gpurun_out/ww_$w.err
python - <<PY
import json,re
d=json.loads(open("gpurun_out/ww_$w.json").read().strip().splitlines()[-1])
waits=re.findall(r"rank (\d): \d+ finalize launches, ([\d.]+) us", open("gpurun_out/ww_$w.err").read())
print("weight $w N=2: Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "sor", round(d["sor"]["ms_per_tick"],3), "pass ms", round(d["roofline"]["avg_launch_ms"],4), "finalize us per rank", sorted(waits))
PY
done
