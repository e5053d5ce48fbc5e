#!/bin/bash
# usage (under gpurun): scripts/gpu_plan_ab.sh -- stream-plan A/B at 8192^2 T=4: tile-granular without walls (old default) vs row-granular with walls
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "stream or walls or red_black" > gpurun_out/pytest_stream.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_stream.log
run() { # name, env...
  name=$1; shift
  env "$@" SB_DEBUG_PLAN=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_plan_$name.json 2> gpurun_out/bench_plan_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_plan_$name.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("$name", "Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "pass ms", round(r["avg_launch_ms"],4), "plan", d["config"]["rb_plan"])
PY
  grep "sb plan" gpurun_out/bench_plan_$name.err | tail -1
}
run old SB_RB_ROW_PLAN=0 SB_RB_STREAM_KINDS=2
run new17 SB_WALL_WEIGHT=1.7
run new15 SB_WALL_WEIGHT=1.5
run new20 SB_WALL_WEIGHT=2.0
run rows_nowalls SB_RB_STREAM_KINDS=2
run old_again SB_RB_ROW_PLAN=0 SB_RB_STREAM_KINDS=2
