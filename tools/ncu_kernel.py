#!/usr/bin/env python
"""print selected `ncu --page raw` metrics of the first launch whose name contains NAME:  ncu_kernel.py REPORT NAME [metric-substring ...]"""
import csv, subprocess, sys
rep, name, subs = sys.argv[1], sys.argv[2], sys.argv[3:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
default = ["gpu__time_duration.sum", "dram__throughput.avg.pct", "sm__warps_active.avg.pct", "smsp__inst_executed.sum", "registers_per_thread",
           "issue_active.avg.pct", "grid_size", "long_scoreboard", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "achieved_occupancy", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "stalled"]
subs = subs or default
for r in rows[2:]:
    if name in r[hdr.index("Kernel Name")]:
        for i, h in enumerate(hdr):
            if any(s in h for s in subs):
                print(f"{h:90s} {r[i]}")
        break
