#!/usr/bin/env python
"""Prints the key metrics of every launch in an `ncu --page raw --csv` export."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "blk"),
        ("launch__occupancy_limit_registers", "occR"), ("launch__occupancy_limit_shared_mem", "occS"),
        ("smsp__inst_executed.sum", "winst")]
print(" | ".join(["kernel"] + [c[1] for c in cols]))
for r in data:
    name = r[idx["Kernel Name"]].replace("void ", "").replace("fno::<unnamed>::", "").split("(")[0][:44]
    vals = []
    for c, _ in cols:
        if c not in idx:
            vals.append("-")
            continue
        v = r[idx[c]].replace(",", "")
        try:
            f = float(v)
            if units[idx[c]] == "byte":
                f /= 1e6
            elif units[idx[c]] == "Kbyte":
                f /= 1e3
            elif units[idx[c]] == "Gbyte":
                f *= 1e3
            elif units[idx[c]] == "ns":
                f /= 1e3
            elif units[idx[c]] == "ms":
                f *= 1e3
            vals.append(f"{f:.1f}" if f != int(f) else f"{int(f)}")
        except ValueError:
            vals.append(v)
    print(" | ".join([name] + vals))
