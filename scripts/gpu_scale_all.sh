#!/bin/bash
# usage (under gpurun --gpus 8): scripts/gpu_scale_all.sh -- weak scaling 1/2/4/8 on one box, then 32768^2 on 8
scripts/gpu_scale.sh "1 2 4 8"
mkdir -p gpurun_out/scale_weak; cp gpurun_out/scale_n*.json gpurun_out/scale_weak/
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --size 32768 32768 > gpurun_out/scale_n8_32768sq.json 2> gpurun_out/scale_n8_32768sq.err
echo "32768^2 rc=$?"; tail -1 gpurun_out/scale_n8_32768sq.json | cut -c1-400; tail -2 gpurun_out/scale_n8_32768sq.err
timeout 600 python -m pytest tests/test_gpu_slabs.py -x -q -m gpu -k torchrun > gpurun_out/pytest_slabs_2gpu.log 2>&1; tail -3 gpurun_out/pytest_slabs_2gpu.log
