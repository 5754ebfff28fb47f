#!/usr/bin/env python
"""Shared-memory wavefronts by instruction class from an `ncu --page source --csv` export (needs --import-source on):
which accesses exceed their ideal wavefront count — the exp table's data-dependent lookups or the [k][tid] rows?
  ncu -i rep.ncu-rep --page source --csv | python profiles/shared_wavefronts.py /dev/stdin"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Address Space" in r)
ix = {h: i for i, h in enumerate(hdr)}
groups, tot = {}, [0.0, 0.0, 0.0]
for r in rows[rows.index(hdr) + 1:]:
    if len(r) < len(hdr) or "hared" not in r[ix["Address Space"]]:
        continue
    w, idl, ex = (float(r[ix[k]] or 0) for k in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Wavefronts Shared Excessive"))
    src = r[ix["Source"]].strip()
    addr = src.split("[")[-1] if "[" in src else ""
    kind = "cp.async (LDGSTS) arrivals" if "LDGSTS" in src else ("exp-table lookup (data-dependent register address, no immediate offset)" if ("LDS" in src and "0x" not in addr)
                                                                  else "rows [k][tid] / staged data (base + immediate offset)")
    g = groups.setdefault((kind, r[ix["Access Operation"]], r[ix["Access Size"]]), [0, 0.0, 0.0, 0.0])
    g[0] += 1; g[1] += w; g[2] += idl; g[3] += ex
    tot[0] += w; tot[1] += idl; tot[2] += ex
print(rows[0][1] if len(rows[0]) > 1 else "")
print("shared wavefronts %.4g, ideal %.4g, excessive %.4g (%.1f %%)" % (tot[0], tot[1], tot[2], 100 * tot[2] / max(tot[0], 1)))
for k, v in sorted(groups.items(), key=lambda x: -x[1][1]):
    print("%-72s %-11s %4s-bit  instructions %3d  wavefronts %.4g  ideal %.4g  ratio %.2f" % (k[0], k[1], k[2], v[0], v[1], v[2], v[1] / max(v[2], 1)))
top = sorted([r for r in rows[rows.index(hdr) + 1:] if len(r) >= len(hdr) and "hared" in r[ix["Address Space"]]],
             key=lambda r: -float(r[ix["L1 Wavefronts Shared Excessive"]] or 0))[:14]
print("largest excess (SASS, wavefronts, ideal, instructions executed):")
for r in top:
    print("  %-58s %12s %12s %12s" % (r[ix["Source"]].strip()[:58], r[ix["L1 Wavefronts Shared"]], r[ix["L1 Wavefronts Shared Ideal"]], r[ix["Instructions Executed"]]))
