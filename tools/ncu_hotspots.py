#!/usr/bin/env python
"""Source-level view of an .ncu-rep captured with `--set full --import-source on` (needs `ncu` on PATH, no GPU):
warp-state samples by stall reason, by opcode, the hottest SASS lines, and samples / instructions per block of
N consecutive SASS lines (to split a kernel into its phases).
usage: tools/ncu_hotspots.py <report.ncu-rep> [--top 25] [--block 64]"""
import csv, io, subprocess, sys
from collections import Counter

STALLS = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_sleep", "stall_math", "stall_selected", "stall_barrier", "stall_mio",
          "stall_lg", "stall_dispatch", "stall_branch_resolving", "stall_no_inst", "stall_not_selected", "stall_tex", "stall_drain", "stall_membar"]


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    block = int(sys.argv[sys.argv.index("--block") + 1]) if "--block" in sys.argv else 64
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][1] if len(rows[0]) > 1 else rows[0][0])
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (KeyError, ValueError, IndexError):
            return 0.0

    samples = sum(f(r, "# Samples") for r in data)
    instrs = sum(f(r, "Instructions Executed") for r in data)
    print(f"{len(data)} SASS lines, {int(samples)} samples, {int(instrs)} warp instructions executed\n")
    print("stall reasons (samples):")
    for k in STALLS:
        v = sum(f(r, k) for r in data)
        if v:
            print(f"  {k:26s} {int(v):8d}  {100 * v / max(samples, 1):5.1f} %")
    by_op, ex_op = Counter(), Counter()
    for r in data:
        words = [w for w in r[ix["Source"]].split() if not w.startswith("@")]
        op = ".".join(words[0].split(".")[:2]) if words else "?"
        by_op[op] += f(r, "# Samples")
        ex_op[op] += f(r, "Instructions Executed")
    print("\nby opcode: samples, instructions executed")
    for op, v in by_op.most_common(top):
        print(f"  {op:22s} {int(v):8d} {int(ex_op[op]):12d}")
    print(f"\nhottest {top} lines: address, samples (long_sb / wait / short_sb / sleep / math), source")
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top]:
        print(f"  {r[ix['Address']][-6:]} {int(f(r, '# Samples')):7d} ({int(f(r, 'stall_long_sb'))}/{int(f(r, 'stall_wait'))}/{int(f(r, 'stall_short_sb'))}/"
              f"{int(f(r, 'stall_sleep'))}/{int(f(r, 'stall_math'))})  {r[ix['Source']][:90]}")
    print(f"\nper block of {block} SASS lines: first line index, samples, instructions, first instruction")
    for b in range(0, len(data), block):
        chunk = data[b : b + block]
        s, i = sum(f(r, "# Samples") for r in chunk), sum(f(r, "Instructions Executed") for r in chunk)
        if s or i:
            print(f"  {b:6d} {int(s):8d} {int(i):12d}  {chunk[0][ix['Source']][:60]}")


main()
