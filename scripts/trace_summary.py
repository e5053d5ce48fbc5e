#!/usr/bin/env python3
"""usage: trace_summary.py <stderr of a run with SB_STREAM_TRACE=1>  -- per-CTA durations of the last streaming pass"""
import collections
import re
import sys

rows = []
for l in open(sys.argv[1]):
    m = re.match(r"\[sb trace\] (?:rank \d+ )?item (\d+) kind (\d) flags (\d+) rows (\d+) sm (\d+) start ([\d.]+) us dur ([\d.]+) us", l)
    if m:
        rows.append(tuple(float(x) for x in m.groups()))
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cta = collections.defaultdict(list)
for r in rows:
    cta[int(r[0]) // nw].append(r)
slow = []
durs = sorted(max(r[6] for r in v) for v in cta.values())
med = durs[len(durs) // 2]
print(f"{len(cta)} CTAs, median CTA duration {med:.0f} us, max {durs[-1]:.0f} us, kernel end {max(r[5] + r[6] for r in rows):.0f} us")
for c, v in sorted(cta.items()):
    d = max(r[6] for r in v)
    if d > 1.15 * med:
        slow.append((c, int(v[0][4]), sorted({int(r[1]) for r in v}), sorted({int(r[2]) for r in v}), int(v[0][3]), round(d)))
print("slow CTAs (cta, sm, kinds, flags, rows, us):", slow)
