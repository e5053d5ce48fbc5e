#!/bin/bash
# usage: scripts/kernel_times.sh [bench args...]   -- per-kernel average durations of one tick (ncu, cold)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_small.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-verify "$@" > /dev/null 2>&1
python - <<EOF
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_small.csv")) if len(r)>10 and r[0].isdigit()]
d=collections.defaultdict(list)
for r in rows: d[r[4].split("(")[0].split("::")[-1]].append(int(r[-1]))
for k,v in d.items(): print(f"{k:28s} n={len(v):4d} avg {sum(v)/len(v)/1e3:9.1f} us  total {sum(v)/1e6:8.3f} ms")
EOF
