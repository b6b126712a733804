#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv` launch list as a markdown table:
per kernel: launches, mean microseconds, share of the summed kernel time.
usage: tools/launch_summary.py launches.csv "title" > profiles/rN_launches_*.md"""
import csv, sys
from collections import defaultdict

def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    t = defaultdict(list)
    for r in rows[1:]:
        name = r[k].split("(")[0].split("::")[-1]
        t[name].append(float(r[v].replace(",", "")))
    total = sum(sum(x) for x in t.values())
    print(f"# {title}\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none --csv`; serialised, cold-cache: compare SHARES. Raw CSV alongside.\n")
    print("| kernel | launches | mean us | share |\n|---|---:|---:|---:|")
    for name, x in sorted(t.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {name} | {len(x)} | {sum(x) / len(x) / 1e3:.1f} | {100 * sum(x) / total:.1f}% |")

main()
