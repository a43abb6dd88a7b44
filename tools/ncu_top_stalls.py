#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel instance from `ncu -i X.ncu-rep --page source --csv` (developer tool).
usage: ncu_top_stalls.py src.csv <section-index> [top-n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
k = int(sys.argv[2]); topn = int(sys.argv[3]) if len(sys.argv) > 3 else 14
a = starts[k]; b = starts[k + 1] if k + 1 < len(starts) else len(rows)
hdr = rows[a + 1]; body = [r for r in rows[a + 2:b] if len(r) >= len(hdr) - 2]
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
print(rows[a][1][:60], "instructions", len(body), "samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for idx, r in sorted(enumerate(body), key=lambda x: -int(x[1][ci["# Samples"]] or 0))[:topn]:
    n = int(r[ci["# Samples"]] or 0)
    top = sorted(((int(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{idx:5d} {100*n/max(tot,1):5.1f}%  {r[1].strip()[:70]:70s} {top}")
