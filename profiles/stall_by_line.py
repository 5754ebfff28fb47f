#!/usr/bin/env python
"""Where the warp-stall samples of a kernel fall, from an `ncu --page source --csv` export (needs --import-source on and
-lineinfo):  ncu -i rep.ncu-rep --page source --csv [--print-source cuda] > src.csv; python profiles/stall_by_line.py src.csv [N]
Prints the sample totals per stall reason and the N rows (SASS instructions or CUDA lines) with the most samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = next(r for r in rows if "Source" in r and any("Samples" in c for c in r))
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[rows.index(hdr) + 1:] if len(r) >= len(hdr)]
def f(r, k):
    try:
        return float(r[ix[k]] or 0)
    except (ValueError, KeyError):
        return 0.0
samp = next((h for h in hdr if h.startswith("# Samples") or h == "Warp Stall Sampling (All Samples)"), None)
stall_cols = [h for h in hdr if h.startswith("stall_")]
print("columns:", [h for h in hdr if h not in stall_cols][:14])
tot = sum(f(r, samp) for r in body)
print("total samples %.0f" % tot)
for h in sorted(stall_cols, key=lambda h: -sum(f(r, h) for r in body))[:10]:
    print("  %-28s %6.2f %%" % (h, 100 * sum(f(r, h) for r in body) / max(tot, 1)))
key = "stall_long_sb" if "stall_long_sb" in ix else samp
for title, k in (("most samples", samp), ("most long-scoreboard samples", key)):
    print("--- " + title)
    for r in sorted(body, key=lambda r: -f(r, k))[:top_n]:
        main = sorted(stall_cols, key=lambda h: -f(r, h))[:2]
        print("  %6.2f %%  %-70s  %s" % (100 * f(r, k) / max(tot, 1), r[ix["Source"]].strip()[:70],
                                         ", ".join("%s %.0f" % (m[6:], f(r, m)) for m in main)))
