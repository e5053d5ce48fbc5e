#!/bin/bash
# usage (under gpurun --gpus 2): scripts/gpu_fin_trace.sh -- slab-mode finalize timing per rank
mkdir -p gpurun_out
SB_FIN_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/fin_trace_n2.json 2> gpurun_out/fin_trace_n2.err
grep "finalize trace" gpurun_out/fin_trace_n2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/fin_trace_n2.json").read().strip().splitlines()[-1])
print("N=2 Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "sor", round(d["sor"]["ms_per_tick"],3), "pass ms", round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["launch_ms_min_median_max"])
PY
