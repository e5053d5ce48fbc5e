#!/bin/bash
# usage (under gpurun): scripts/gpu_confirm.sh -- C3, C4 and the default line after a plan change
mkdir -p gpurun_out
for w in c3 c4; do
timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
print("$w", d["config"]["grid"], "Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "pass", round(d["roofline"]["avg_launch_ms"],4), d["config"]["rb_plan"])
PY
done
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-330 gpurun_out/bench_default.json
