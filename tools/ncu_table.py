#!/usr/bin/env python3
"""One markdown table row per captured launch from `ncu -i report.ncu-rep --page raw --csv` output.
usage: ncu_table.py raw.csv > table.md"""
import csv
import re
import sys

COLS = [("gpu__time_duration.sum", "time", lambda v, u: f"{float(v) / (1000.0 if u == 'ns' else 1.0):.1f} us" if u in ("ns", "us") else f"{v} {u}"),
        ("launch__registers_per_thread", "regs", lambda v, u: f"{float(v):.0f}"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %", lambda v, u: f"{float(v):.1f}"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", lambda v, u: f"{float(v):.1f}"),
        ("smsp__inst_executed.sum", "warp inst", lambda v, u: f"{float(v) / 1e6:.2f} M"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %", lambda v, u: f"{float(v):.1f}"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %", lambda v, u: f"{float(v):.1f}"),
        ("dram__bytes_read.sum", "dram rd", None), ("dram__bytes_write.sum", "dram wr", None),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 %", lambda v, u: f"{float(v):.1f}"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", lambda v, u: f"{float(v):.1f}")]


def mb(v, u):
    f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    return f"{float(v) * f:.3f} MB"


rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr, units = rows[0], rows[1]
print("| # | kernel | grid | block | " + " | ".join(c[1] for c in COLS) + " |")
print("|---|---|---|---|" + "---:|" * len(COLS))
for i, r in enumerate(rows[2:]):
    g = lambda k: (r[hdr.index(k)], units[hdr.index(k)]) if k in hdr else ("0", "")
    name = re.sub(r"\(.*", "", g("Kernel Name")[0]).replace("void ", "")
    cells = []
    for key, _, fmt in COLS:
        v, u = g(key)
        v = v.replace(",", "")
        try:
            cells.append(mb(v, u) if fmt is None else fmt(v, u))
        except ValueError:
            cells.append(v)
    print(f"| {i} | `{name}` | {g('Grid Size')[0]} | {g('Block Size')[0]} | " + " | ".join(cells) + " |")
