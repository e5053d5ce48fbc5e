#!/bin/bash
# usage (under gpurun --gpus 2): scripts/gpu_swap.sh -- is the slower slab a slower GPU or a slower rank? (devices swapped in the second run)
mkdir -p gpurun_out
for vis in 0,1 1,0; do
CUDA_VISIBLE_DEVICES=$vis SB_FIN_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/swap_$vis.json 2> gpurun_out/swap_$vis.err
echo "devices $vis"; grep "finalize trace" gpurun_out/swap_$vis.err | cut -c1-140
python - <<PY
import json
d=json.loads(open("gpurun_out/swap_$vis.json").read().strip().splitlines()[-1])
print("  Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "pass ms", round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["launch_ms_min_median_max"])
PY
done
for dev in 0 1; do
CUDA_VISIBLE_DEVICES=$dev timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/swap_single_$dev.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/swap_single_$dev.json").read().strip().splitlines()[-1])
print("single GPU $dev: Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "pass ms", round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["launch_ms_min_median_max"])
PY
done
