#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python profiles/launch_table.py file.csv [-v]"""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
tot = {}
for r in rows[1:]:
    name = re.sub(r"^void ", "", r[ki])
    m = re.match(r"(cude::)?([a-z_0-9]+)(<.*>)?", name)
    short = m.group(2) + ((" " + m.group(3)[:70]) if m and m.group(3) else "") if m else name[:80]
    ms = float(r[vi].replace(",", "")) / 1e6
    if "-v" in sys.argv:
        print(f"{short:100s} grid {r[gi]:>16s} block {r[bi]:>12s} {ms:9.3f} ms")
    t = tot.setdefault(short, [0, 0.0]); t[0] += 1; t[1] += ms
all_ms = sum(v[1] for v in tot.values())
for k, v in sorted(tot.items(), key=lambda x: -x[1][1]):
    print(f"{k:100s} n={v[0]:4d} total {v[1]:9.3f} ms  {100 * v[1] / all_ms:5.1f} %")
print(f"{'all':100s}        total {all_ms:9.3f} ms")
