#!/bin/bash
# usage (under gpurun): scripts/gpu_ncu_mid.sh -- ncu --set full of the grid-resident solve on config 2 + source-level stall summary
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sor_mid_reg -s 1 -c 1 -o gpurun_out/mid_reg_c2_full -f python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu > gpurun_out/mid_reg_c2_full.log 2>&1
tail -3 gpurun_out/mid_reg_c2_full.log
ls -la gpurun_out/*.ncu-rep
