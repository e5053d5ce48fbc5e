#!/bin/bash
# usage: scripts/bench_variants.sh "<lib paths>" "<T list>"
for lib in $1; do for T in $2; do
SB_LIB=$lib timeout 300 python bench.py --steps 3 --warmup 3 --tblock $T --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', d['config']['temporal_block'], round(d['value'],1), round(d['ms_per_step'],2), round(d['roofline']['avg_launch_ms'],4), round(d['roofline']['frac'],3))"
done; done
