#!/bin/bash
# usage (under gpurun --gpus 2): scripts/gpu_trace4.sh -- wall items per row vs the row pitch pad (SB_PITCH_PAD, elements), both slabs
mkdir -p gpurun_out
for pad in 0 64 272 1040; do
SB_PITCH_PAD=$pad SB_STREAM_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/trace4_$pad.json 2> gpurun_out/trace4_$pad.err
python - <<PY
import re,collections,json
d=json.loads(open("gpurun_out/trace4_$pad.json").read().strip().splitlines()[-1])
print("pad $pad: ms/step", round(d["ms_per_step"],3), "pass", round(d["roofline"]["avg_launch_ms"],4))
rows=collections.defaultdict(list)
for l in open("gpurun_out/trace4_$pad.err"):
    m=re.match(r"\[sb trace\] rank (\d+) item (\d+) kind (\d) flags (\d+) rows (\d+) sm (\d+) start ([\d.]+) us dur ([\d.]+) us warmup ([\d.]+) steady ([\d.]+) drain ([\d.]+)", l)
    if m: rows[int(m.group(1))].append([float(x) for x in m.groups()[1:]])
for r,v in sorted(rows.items()):
    end=max(x[5]+x[6] for x in v)
    out=[f"rank {r} span {end:6.1f} us"]
    for kind in (0,1,2):
        g=[x for x in v if int(x[1])==kind and int(x[2])==kind]
        if g:
            n=len(g); avg=lambda i: sum(x[i] for x in g)/n
            out.append(f"kind {kind}: n={n} rows {avg(3):.0f} dur {avg(6):.1f} steady {avg(8):.1f} ({avg(8)/max(avg(3)-30,1):.3f} us/row)")
    print("  "+" | ".join(out))
PY
done
