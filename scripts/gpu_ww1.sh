#!/bin/bash
# usage (under gpurun --gpus 2): scripts/gpu_ww1.sh -- wall weight sweep on 1 GPU, then more at 2 slabs
mkdir -p gpurun_out
for w in 2.0 3.0 4.0 5.0 6.0; do
SB_WALL_WEIGHT=$w timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/ww1_$w.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/ww1_$w.json").read().strip().splitlines()[-1])
print("weight $w N=1: Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "pass ms", round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["launch_ms_min_median_max"])
PY
done
scripts/gpu_ww.sh "2.0 5.0 6.0 8.0"
