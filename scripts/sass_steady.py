#!/usr/bin/env python3
"""usage: sass_steady.py <object>  -- the steady-state loop of each streaming kernel: size, spills, mix"""
import collections
import re
import subprocess
import sys

txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
ins = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)([^;]*);")
for f in txt.split("Function :")[1:]:
    name = f.split("\n")[0]
    m = re.search(r"stream_kernelILi(\d)", name)
    if not m:
        continue
    lines = [(int(mm.group(1), 16), mm.group(3), mm.group(4)) for mm in ins.finditer(f)]
    for a, op, rest in lines:
        if op != "BRA":
            continue
        mm = re.search(r"0x([0-9a-f]+)", rest)
        if mm and int(mm.group(1), 16) < a:
            tgt = int(mm.group(1), 16)
            n = (a - tgt) // 16 + 1
            ops = collections.Counter(o for ad, o, _ in lines if tgt <= ad <= a)
            if n > 500 and ops["ISETP"] < 12:  # straight-line body without row tests
                nw = 2 * int(m.group(1)) + 4
                print(f"T={m.group(1)}: {n} instr / {nw} ticks = {n / nw:.0f} per tick, "
                      f"{n * 16 / 1024:.1f} KB, LDL {ops['LDL']} STL {ops['STL']}", dict(ops.most_common(16)))
                break
