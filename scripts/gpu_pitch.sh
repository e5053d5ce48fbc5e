#!/bin/bash
# usage (under gpurun --gpus 2): scripts/gpu_pitch.sh -- row pitch padded by 128 B vs power-of-two pitch: tests, 1 GPU, 2 slabs with the per-item trace
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
for pad in 16 0 32; do
SB_PITCH_PAD=$pad timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/pitch_n1_$pad.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/pitch_n1_$pad.json").read().strip().splitlines()[-1])
print("pad $pad single GPU: Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "sor", round(d["sor"]["ms_per_tick"],3), "pass ms", round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["launch_ms_min_median_max"], "e2e", round(d["e2e"]["value"],1))
PY
done
for pad in 16 0; do
SB_PITCH_PAD=$pad SB_FIN_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/pitch_n2_$pad.json 2> gpurun_out/pitch_n2_$pad.err
grep "finalize trace" gpurun_out/pitch_n2_$pad.err | cut -c1-130
python - <<PY
import json
d=json.loads(open("gpurun_out/pitch_n2_$pad.json").read().strip().splitlines()[-1])
print("pad $pad N=2: Mcs/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "sor", round(d["sor"]["ms_per_tick"],3), "pass ms", round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["launch_ms_min_median_max"])
PY
done
