#!/usr/bin/env python
"""Launch list -> per-kernel table.  Input: the CSV of
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c N --csv python bench.py ...
    python tools/ncu_launch_summary.py launches.csv <training steps seen by the capture> > table.md"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
per = defaultdict(lambda: defaultdict(float))
ids = defaultdict(set)
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    name = r[ix["Kernel Name"]]
    name = name.split("(")[0].replace("void ", "").replace("fno::<unnamed>::", "").replace("unnamed>::", "").strip()
    val = float(r[ix["Metric Value"]].replace(",", "") or 0)
    unit = r[ix["Metric Unit"]]
    m = r[ix["Metric Name"]]
    if m == "gpu__time_duration.sum":
        val = val / 1000.0 if unit in ("ns", "nsecond") else (val * 1000.0 if unit in ("ms", "msecond") else val)
    else:
        val = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6) * val
    per[name][m] += val
    ids[name].add(r[ix["ID"]])
tot = sum(v["gpu__time_duration.sum"] for v in per.values())
print("| kernel | launches / step | avg µs | µs / step | share | DRAM MB / launch (read + write) |")
print("|---|---:|---:|---:|---:|---:|")
for name, v in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    n = len(ids[name])
    t = v["gpu__time_duration.sum"]
    if t / tot < 0.001:
        continue
    print(f"| `{name[:64]}` | {n / steps:.1f} | {t / n:.1f} | {t / steps:.1f} | {100 * t / tot:.1f} % | "
          f"{v['dram__bytes_read.sum'] / n:.1f} + {v['dram__bytes_write.sum'] / n:.1f} |")
print(f"| total | | | {tot / steps:.1f} | | |")
