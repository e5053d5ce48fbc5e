#!/bin/bash
# usage: scripts/all_workloads.sh [extra bench args]  -- one bench line per BASELINE config at N=1
for w in c1 c2 c3 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-verify "$@" > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
    r=d["roofline"] or {}
    print("$w", d["config"]["grid"], "T", d["config"]["temporal_block"], "Mcs/s", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "sweeps/tick", d["config"]["sweeps_per_tick"], "pass ms", round(r.get("avg_launch_ms",0),4), "frac", round(r.get("frac",0),3), "launches", d["gpu_launches"], "e2e", round(d["e2e"]["value"],1))
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/bench_$w.err").read()[-1500:])
PY
done
