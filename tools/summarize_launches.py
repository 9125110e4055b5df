#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share.
usage: summarize_launches.py launches.csv [out.md]"""
import csv
import re
import sys
from collections import defaultdict


def main(path, out=None):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        short = re.sub(r"cub::CUB_\d+_SM_\d+::", "cub::", short)
        short = re.sub(r"<.*", "<...>", short)
        rows.append((short, float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for n, ns, _, _ in rows:
        a = agg[n]
        a[0] += 1; a[1] += ns; a[2] = max(a[2], ns)
    total = sum(a[1] for a in agg.values())
    lines = [f"launches: {len(rows)}   total device time: {total / 1e6:.3f} ms (serialised, cold cache: compare SHARES)", "",
             "| kernel | launches | total ms | share | avg us | max us |", "|---|---:|---:|---:|---:|---:|"]
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{n}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / total:.1f}% | {a[1] / a[0] / 1e3:.2f} | {a[2] / 1e3:.1f} |")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "a").write(text)
    print(text)


if __name__ == "__main__":
    main(*sys.argv[1:3])
