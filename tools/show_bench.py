#!/usr/bin/env python
"""Pretty-print the per-kernel table of a bench.py JSON line (developer tool)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print("value %.3f G gates/s  %.3f ms/step  e2e %.1f M gates/s (%.2f ms)  launches %s clocks %s" % (
    d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["s_per_step"] * 1e3, d["gpu_launches"], d["clocks"]))
print("dominant %s frac %.3f  whole-step %.0f GB/s" % (r["kernel"], r["frac"], r["whole_step_gbs"]))
for k, v in sorted(r["per_kernel_ms"].items(), key=lambda x: -x[1]):
    print(f"  {k:28s} {v:.4f} ms  {r['per_kernel_gbs'].get(k)}")
print("  kernel sum %.3f ms" % sum(r["per_kernel_ms"].values()))
