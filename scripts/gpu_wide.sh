#!/bin/bash
# usage (under gpurun): scripts/gpu_wide.sh -- the per-slab plan of 32768^2 on 8 GPUs, emulated on one GPU (4096 x 32768)
mkdir -p gpurun_out
for v in 1 0; do
SB_RB_ROW_PLAN=$v SB_DEBUG_PLAN=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --size 4096 32768 > gpurun_out/wide_$v.json 2> gpurun_out/wide_$v.err
grep "sb plan" gpurun_out/wide_$v.err | tail -1
python - <<PY
import json
d=json.loads(open("gpurun_out/wide_$v.json").read().strip().splitlines()[-1])
print("row plan=$v 4096x32768: Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "pass ms", round(d["roofline"]["avg_launch_ms"],4), d["config"]["rb_plan"])
PY
done
timeout 600 python -m pytest tests -m gpu -x -q -k "stream or walls or red_black or slab" > gpurun_out/pytest_stream.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_stream.log
