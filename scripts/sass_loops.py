#!/usr/bin/env python3
"""usage: sass_loops.py <object or .so> [name filter]  -- per kernel: instruction count, the
largest loops (backward branches) and the opcode mix of each (steady-state cost per iteration)."""
import collections
import re
import subprocess
import sys

txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
flt = sys.argv[2] if len(sys.argv) > 2 else ""
ins = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)([^;]*);")
for f in txt.split("Function :")[1:]:
    name = f.split("\n")[0].strip()
    if flt not in name:
        continue
    lines = [(int(m.group(1), 16), m.group(3), m.group(4)) for m in ins.finditer(f)]
    loops = []
    for a, op, rest in lines:
        if op == "BRA":
            mm = re.search(r"0x([0-9a-f]+)", rest)
            if mm and int(mm.group(1), 16) < a:
                loops.append((a - int(mm.group(1), 16), int(mm.group(1), 16), a))
    loops.sort(reverse=True)
    print(name[-90:], "instructions:", len(lines))
    for size, tgt, a in loops[:3]:
        ops = collections.Counter(op for ad, op, _ in lines if tgt <= ad <= a)
        print(f"  loop {size // 16 + 1} instr @{tgt:#x}:", dict(ops.most_common(24)))
