#!/usr/bin/env python
"""per-source-line executed warp instructions of a kernel in an ncu report (needs -lineinfo + --import-source on):
   ncu_lines.py REPORT KERNEL-SUBSTRING [top]"""
import csv, subprocess, sys
rep, name = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fn, fpath, hdr, acc = None, None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "Function Name":
        fn = r[1]
    elif r[0] == "File Path":
        fpath = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and fn and name in fn and r[0].isdigit():
        i_exec = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        key = (fn[:60], fpath.split("/")[-1], int(r[0]))
        try:
            v = (int(r[i_exec]), int(r[i_samp]), r[1].strip()[:110])
        except ValueError:
            continue
        old = acc.get(key)
        acc[key] = (v[0] + (old[0] if old else 0), v[1] + (old[1] if old else 0), v[2])
fns = sorted({k[0] for k in acc})
for f in fns:
    items = [(k, v) for k, v in acc.items() if k[0] == f]
    tot = sum(v[0] for _, v in items)
    if not tot:
        continue
    print(f"== {f}: {tot / 1e6:.1f} M warp instructions")
    for k, v in sorted(items, key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0] / 1e6:8.2f} M {100 * v[0] / tot:5.1f}%  smp {v[1]:5d}  {k[1]}:{k[2]:<5d} {v[2]}")
