#!/bin/bash
# usage (under gpurun): scripts/gpu_ncu.sh <kernel regex> <out name> [bench args]
# one `ncu --set full` capture of the 10th matching launch -> gpurun_out/<out name>.ncu-rep
K=$1; O=$2; shift 2
ncu --set full --clock-control none --import-source on -k regex:$K -s 10 -c 1 -o gpurun_out/$O -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-verify "$@" > gpurun_out/$O.log 2>&1
tail -3 gpurun_out/$O.log
