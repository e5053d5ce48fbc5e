#!/usr/bin/env python3
"""Digest of the source page of an `ncu --set full --import-source on` report (works without a GPU):

    python scripts/ncu_source_digest.py gpurun_out/x.ncu-rep > profiles/<name>.txt

Prints the stall-reason totals of the launch, the share of warp-stall samples per opcode and the
25 SASS instructions with the most samples.
"""
import collections
import csv
import io
import subprocess
import sys


def page(rep, what, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", what, "--csv", *extra], check=True,
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, row = raw[0], raw[2]
    print(f"# {rep}: kernel {row[hdr.index('Kernel Name')][:100]}")
    stalls = []
    for i, k in enumerate(hdr):
        if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
            try:
                stalls.append((float(row[i]), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in stalls) or 1.0
    print("## warp-stall samples by reason")
    for v, k in sorted(stalls, reverse=True)[:12]:
        print(f"  {k:28s} {v:10.0f}  {100 * v / tot:5.1f} %")
    rows = page(rep, "source")
    h = next(i for i, r in enumerate(rows) if "# Samples" in r)
    col = {k: i for i, k in enumerate(rows[h])}
    data = [r for r in rows[h + 1:] if len(r) > col["# Samples"] and r[col["# Samples"]].isdigit()]
    total = sum(int(r[col["# Samples"]]) for r in data) or 1
    ops = collections.Counter()
    for r in data:
        op = r[col["Source"]].strip().split()
        op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "?")
        ops[op.split(".")[0]] += int(r[col["# Samples"]])
    print(f"## samples by opcode they are attributed to ({total} samples, {len(data)} instructions)")
    for op, n in ops.most_common(12):
        print(f"  {100 * n / total:5.1f} %  {op}")
    print("## the 25 SASS instructions with the most samples")
    for r in sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:25]:
        n = int(r[col["# Samples"]])
        print(f"  {100 * n / total:5.1f} %  {r[col['Address']][-5:]}  {r[col['Source']].strip()[:100]}")


if __name__ == "__main__":
    main()
