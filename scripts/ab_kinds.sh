#!/bin/bash
# usage: scripts/ab_kinds.sh "<T list>" "<kinds list>" -- A/B on one box: committed library (variants/lib_head.so) vs working tree
run() { python bench.py --steps 4 --warmup 2 --no-cpu --tblock $T 2>&1 | grep -E "sb plan|metric" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('[sb'): print('   ',l.strip())
    else:
        d=json.loads(l); print('    Mcs/s',round(d['value'],1),'pass ms',round(d['roofline']['avg_launch_ms'],4),'sor ms',round(d['sor']['ms_per_tick'],2))
"; }
for T in $1; do
echo "T=$T HEAD lib"; SB_LIB=stroemung_b200/variants/lib_head.so run
for K in $2; do echo "T=$T working tree KINDS=$K"; SB_RB_STREAM_KINDS=$K SB_DEBUG_PLAN=1 run; done
done
