#!/bin/bash
# usage (under gpurun): scripts/gpu_iter.sh "<T list>" [pytest -k filter]  -- GPU parity tests then bench per T
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
for T in $1; do
  timeout 300 python bench.py --steps 10 --warmup 3 --tblock $T --no-cpu > gpurun_out/bench_T$T.json 2> gpurun_out/bench_T$T.err; echo "bench T=$T rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_T$T.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("T=$T Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "pass ms", round(r["avg_launch_ms"],4), "sor ms", round(d["sor"]["ms_per_tick"],3), "e2e", round(d["e2e"]["value"],1), "clk", d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_T$T.err").read()[-2000:])
PY
done
