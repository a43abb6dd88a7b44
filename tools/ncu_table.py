#!/usr/bin/env python
"""Markdown table of the key metrics of every kernel in an .ncu-rep (via `ncu --page raw --csv`); developer tool.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; ncu_table.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, body = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k, d=0.0):
    try: return float(r[ix[k]].replace(",", ""))
    except Exception: return d
print("| kernel | time us | DRAM read MB | DRAM write MB | DRAM GB/s | dram % | warps active % | regs | grid | L2 hit % |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for r in body:
    name = r[ix["Kernel Name"]].split("(")[0].replace("c2a::", "")
    t = g(r, "gpu__time_duration.sum")
    tu = rows[1][ix["gpu__time_duration.sum"]]
    t_us = t / 1000 if tu == "ns" else (t if tu == "us" else t * 1000)
    def mb(k):
        v = g(r, k); u = rows[1][ix[k]]
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
    rd, wr = mb("dram__bytes_read.sum"), mb("dram__bytes_write.sum")
    print(f"| `{name}` | {t_us:.1f} | {rd:.1f} | {wr:.1f} | {(rd + wr) / t_us * 1e3:.0f} | {g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {int(g(r, 'launch__registers_per_thread'))} | {int(g(r, 'launch__grid_size'))} | {g(r, 'lts__t_sector_hit_rate.pct'):.0f} |")
