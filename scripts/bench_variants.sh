#!/bin/bash
# usage: scripts/bench_variants.sh "<lib paths>" "<T list>" [extra bench args]
L=$1; TS=$2; shift 2
for lib in $L; do for T in $TS; do
SB_LIB=$lib timeout 300 python bench.py --steps 5 --warmup 3 --tblock $T --no-cpu "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'T', d['config']['temporal_block'], 'Mcs/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'pass ms', round(d['roofline']['avg_launch_ms'],4), 'sor ms', round(d['sor']['ms_per_tick'],2))"
done; done
