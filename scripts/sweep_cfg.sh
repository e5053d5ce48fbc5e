#!/bin/bash
# usage: scripts/sweep_cfg.sh "<cluster sizes>" "<T list>"  -- bench value / ms per pass for each
for CL in $1; do for T in $2; do
SB_RB_CLUSTER=$CL timeout 300 python bench.py --steps 3 --warmup 3 --tblock $T --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CL=$CL T=$T', 'Mcs/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'pass ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'sor ms', round(d['sor']['ms_per_tick'],2))"
done; done
