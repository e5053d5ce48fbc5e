#!/bin/bash
# usage (under gpurun): scripts/gpu_frozen.sh -- frozen-tile tests, then config 4 with / without them
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "frozen" > gpurun_out/pytest_frozen.log 2>&1; echo "pytest frozen rc=$?"; tail -12 gpurun_out/pytest_frozen.log
for v in 1 0; do
SB_RB_FROZEN=$v SB_DEBUG_PLAN=1 timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c4_frozen$v.json 2> gpurun_out/bench_c4_frozen$v.err
grep "sb plan" gpurun_out/bench_c4_frozen$v.err | tail -1
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c4_frozen$v.json").read().strip().splitlines()[-1])
print("c4 frozen=$v", "Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "pass", round(d["roofline"]["avg_launch_ms"],4), d["config"]["rb_plan"], "launches", d["gpu_launches"])
PY
done
timeout 900 python -m pytest tests -m gpu -x -q -k "stream or walls or red_black or slab" > gpurun_out/pytest_stream.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_stream.log
