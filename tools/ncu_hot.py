#!/usr/bin/env python
"""Lists the SASS instructions of an `ncu --page source --print-source sass --csv` export that
collected the most warp-stall samples, with their dominant stall reasons.
    python tools/ncu_hot.py k.csv [min_fraction=0.008]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.008
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}


def num(r, h):
    try:
        return int(r[ix[h]])
    except (ValueError, IndexError):
        return 0


data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
tot = sum(num(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data))
for i, r in enumerate(data):
    n = num(r, "# Samples")
    if n > tot * frac:
        st = {h: num(r, h) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(f"{i:5d} {100 * n / tot:5.1f}%  {r[ix['Source']].strip()[:64]:64s} x{num(r, 'Instructions Executed'):<9d} {top}")
