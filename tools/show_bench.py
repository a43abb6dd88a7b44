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
for key in ("same_config",):
    if key in d:
        x = d[key]
        print(f"{key}: {x['gates']} gates value {x['value']/1e6:.1f} M/s ({x['ms_per_step']*1e3:.1f} us) flushed {x['ms_per_step_l2_flushed']*1e3:.1f} us e2e {x['e2e']['value']/1e6:.1f} M/s "
              f"launches {x['gpu_launches_per_step']} cpu_backend {x.get('cpu_backend_gates_per_s', 0)/1e6:.2f} M/s ratio_e2e {x.get('ratio_e2e_vs_reference')}")
for name, x in d.get("configs", {}).items():
    print(f"config {name}: {x['gates']} gates value {x['value']/1e6:.1f} M/s ({x['ms_per_step']*1e3:.1f} us) flushed {x['ms_per_step_l2_flushed']*1e3:.1f} us e2e {x['e2e']['value']/1e6:.1f} M/s "
          f"({x['e2e']['s_per_step']*1e6:.0f} us) conc {x.get('value_concurrent', {}).get('value', 0)/1e6:.1f} M/s sort {x['topo_sort_ms']*1e3:.1f} us {x['topo_hbm_gbs']} GB/s launches {x['gpu_launches_per_step']} cpu_backend {x.get('cpu_backend_gates_per_s', 0)/1e6:.2f} M/s {x.get('parity_vs_oracle')}")
for name, x in d.get("variants", {}).items():
    print(f"variant {name}: {x['gates']} gates value {x['value']/1e6:.1f} M/s ({x['ms_per_step']:.3f} ms) sort {x['topo_sort_ms']:.3f} ms cpu_backend {x.get('cpu_backend_gates_per_s', 0)/1e6:.2f} M/s fallback {x.get('relax_fallback_rounds')}")
    if "phases_ms" in x:
        print("    ", x["phases_ms"])
for name, x in d.get("worst_case_shapes", {}).items():
    print(f"deep {name}: {x['gates']} gates value {x['value']/1e6:.1f} M/s ({x['ms_per_step']:.3f} ms) cpu_backend {x.get('cpu_backend_gates_per_s', 0)/1e6:.2f} M/s x{x.get('speedup_vs_cpu_backend', 0):.1f} rounds {x.get('relax_fallback_rounds')}")
    print("    ", x["phases_ms"])
for name, x in d.get("kahn", {}).items():
    print(f"kahn {name}: {x['gates']} gates {x['levels']} levels kernels {x['ms_kernels']:.3f} ms call {x['ms_call']:.3f} ms {x['achieved_gbs']:.0f} GB/s frac {x['frac']:.3f}")
    print("    ", x["phases_ms"])
if "sweeps" in d:
    print("sweeps", {k: v for k, v in d["sweeps"].items() if k != "note"})
for k in ("strong_scaling", "e2e_pipelined", "from_source", "e2e_host_emitter", "cpu_baseline", "multi_gpu_parity", "extras_wall_s"):
    if k in d:
        print(k, d[k])
