#!/usr/bin/env python
"""Summarises `ncu --page source --csv --print-source sass` output: executed-instruction mix by
opcode and the warp-stall breakdown of one kernel launch.

    ncu -i rep.ncu-rep --page source --csv --print-source sass --launch-skip N --launch-count 1 > k.csv
    python tools/ncu_sass_summary.py k.csv
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
name = rows[0][1] if rows[0][0] == "Kernel Name" else "?"
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
stalls = collections.Counter()
total = 0
samples = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[ix["Instructions Executed"]])
    except ValueError:
        continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "LDG", "STG", "STS", "LDL", "STL")) else op.split(".")[0]
    ops[op] += n
    total += n
    samples += int(r[ix["# Samples"]] or 0)
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            try:
                stalls[h] += int(r[ix[h]])
            except ValueError:
                pass
print(name[:120])
print(f"warp instructions executed: {total}")
for op, n in ops.most_common(24):
    print(f"  {op:14s} {n:12d} {100 * n / total:6.2f}%")
st = sum(stalls.values())
print(f"stall samples: {st}")
for h, n in stalls.most_common(10):
    print(f"  {h:28s} {n:8d} {100 * n / max(st, 1):6.2f}%")
