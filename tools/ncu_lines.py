#!/usr/bin/env python3
"""Hottest CUDA source lines of one kernel in an `ncu --set full --import-source on` report.
usage: ncu_lines.py report.ncu-rep kernel_name [top_n]   (kernel compiled with -lineinfo)"""
import csv
import io
import subprocess
import sys


def main(rep, kernel, top=25):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    file, hdr, lines = "", None, []
    launches = 0
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            file = r[1].split("/")[-1]
        elif len(r) == 2 and r[0] == "Function Name":
            launches += 1
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            g = lambda k: float(r[hdr.index(k)]) if r[hdr.index(k)] not in ("-", "") else 0.0
            lines.append((file, int(r[0]), r[1].strip(), g("Instructions Executed"), g("Thread Instructions Executed"),
                          g("Warp Stall Sampling (All Samples)"), launches))
    merged = {}
    for l in lines:                                        # several captured launches of one kernel: summed
        k = (l[0], l[1])
        m = merged.get(k)
        merged[k] = l if m is None else (l[0], l[1], l[2], m[3] + l[3], m[4] + l[4], m[5] + l[5], 0)
    lines = list(merged.values())
    ti, ts = sum(l[3] for l in lines), sum(l[5] for l in lines)
    print(f"{kernel}: {ti:.3g} warp instructions, {ts:.0f} stall samples")
    print("| file:line | inst % | samples % | threads/inst | source |\n|---|---:|---:|---:|---|")
    for l in sorted(lines, key=lambda l: -l[3])[:top]:
        print(f"| {l[0]}:{l[1]} | {100 * l[3] / ti:.1f} | {100 * l[5] / max(ts, 1):.1f} | {l[4] / max(l[3], 1):.1f} | `{l[2][:110]}` |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
