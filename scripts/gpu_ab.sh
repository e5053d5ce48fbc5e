#!/bin/bash
# usage (under gpurun): scripts/gpu_ab.sh <tag> <lib.so or -> [bench args]  -- per-kernel times of one tick + a bench line
TAG=$1; LIB=$2; shift 2
[ "$LIB" != "-" ] && export SB_LIB=$PWD/$LIB
echo "== $TAG (${SB_LIB:-default lib}) $@"
scripts/kernel_times.sh "$@" 2>&1 | grep -E "stream|fg_rhs|adapt_uv|sor_rb_kernel|finalize|range_k"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-verify "$@" > gpurun_out/ab_$TAG.json 2> gpurun_out/ab_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$TAG.json").read().strip().splitlines()[-1])
    r=d["roofline"] or {}
    print("$TAG", d["config"]["grid"], "Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "sor ms", round(d["sor"]["ms_per_tick"],3), "pass ms", round(r.get("avg_launch_ms",0),4), "stages", {k: round(v,3) for k,v in d["stage_ms_per_tick"].items()}, "launches", d["gpu_launches"])
except Exception as e:
    print("$TAG failed", e); print(open("gpurun_out/ab_$TAG.err").read()[-1500:])
PY
