#!/bin/bash
# usage (under gpurun --gpus 2): scripts/gpu_slab_edge.sh -- slab tests, then 2-slab bench with slab-edge rows in the stream (default) vs on the tile kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slabs.py -m gpu -x -q > gpurun_out/pytest_slabs.log 2>&1; echo "pytest slabs rc=$?"; tail -4 gpurun_out/pytest_slabs.log
for v in 0 1; do
SB_SLAB_EDGE_TILES=$v SB_FIN_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/slab_edge_$v.json 2> gpurun_out/slab_edge_$v.err
grep "finalize trace" gpurun_out/slab_edge_$v.err
python - <<PY
import json
d=json.loads(open("gpurun_out/slab_edge_$v.json").read().strip().splitlines()[-1])
print("edge tiles=$v N=2 Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "sor", round(d["sor"]["ms_per_tick"],3), "pass ms", round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["launch_ms_min_median_max"], d["config"]["rb_plan"])
PY
done
