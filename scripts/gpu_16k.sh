#!/bin/bash
# usage (under gpurun): scripts/gpu_16k.sh -- BASELINE config 5 at 16384^2 on one GPU
mkdir -p gpurun_out
SB_DEBUG_PLAN=1 timeout 240 python bench.py --size 16384 16384 --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_16384sq.json 2> gpurun_out/bench_16384sq.err; echo rc=$?
grep "sb plan" gpurun_out/bench_16384sq.err | tail -1
cut -c1-900 gpurun_out/bench_16384sq.json
