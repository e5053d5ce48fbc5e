#!/bin/bash
# usage (under gpurun --gpus 2): scripts/gpu_trace2.sh -- per-item trace of the last streaming pass on both slabs
mkdir -p gpurun_out
SB_STREAM_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/trace2.json 2> gpurun_out/trace2.err
grep -c "sb trace" gpurun_out/trace2.err
python - <<PY
import re,collections
rows=collections.defaultdict(list)
for l in open("gpurun_out/trace2.err"):
    m=re.match(r"\[sb trace\] rank (\d+) item (\d+) kind (\d) flags (\d+) rows (\d+) sm (\d+) start ([\d.]+) us dur ([\d.]+) us warmup ([\d.]+) steady ([\d.]+) drain ([\d.]+)", l)
    if m: rows[int(m.group(1))].append([float(x) for x in m.groups()[1:]])
for r,v in sorted(rows.items()):
    end=max(x[5]+x[6] for x in v)
    print("rank",r,"items",len(v),"kernel span",round(end,1),"us")
    groups=collections.defaultdict(list)
    for x in v: groups[(int(x[1]),int(x[2]))].append(x)
    for k,g in sorted(groups.items()):
        n=len(g); avg=lambda i: sum(x[i] for x in g)/n
        print(f"   kind {k[0]} flags {k[1]:3d}: n={n:4d} rows {avg(3):6.0f} start {avg(5):6.1f} dur {avg(6):6.1f} (max {max(x[6] for x in g):6.1f}) warmup {avg(7):6.1f} steady {avg(8):6.1f} drain {avg(9):6.1f}")
PY
