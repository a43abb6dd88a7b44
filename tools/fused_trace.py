#!/usr/bin/env python
"""Per-barrier timeline of the fused kernel (C2A_FUSED_TRACE=1): developer tool, GPU only."""
import os, sys
os.environ["C2A_FUSED_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from c2a_loader import c2a
ctx = c2a.DeviceContext(0)
per_cta = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
c2a.lib.c2a_set_fused_limits(1 << 22, per_cta)
for wl in (c2a.workloads.poseidon_shaped(), c2a.workloads.sha256_shaped(), c2a.workloads.keccak_shaped(1), c2a.workloads.mimc_chains(37, 91, "late"), c2a.workloads.mimc_chains(1832, 91, "late")):
    k, w, f = c2a.pack_events(np.ascontiguousarray(wl.events))
    ins, outs = np.array(sorted(wl.inputs), dtype=np.uint32), np.array(sorted(wl.outputs), dtype=np.uint32)
    for _ in range(3):
        ctx.compile_packed(k, w, f, ins, outs, want_order=False, want_wires=False)
    ph = ctx.phases()
    tr = {a: b for a, b in ph.items() if a.startswith("fused:")}
    print(wl.name, "gates", wl.n_gates, "events", len(k), "grid", tr.pop("fused:grid", None), "kernel %.1f us" % (ph["k_fused_compile"] * 1e3), "h2d %.1f d2h %.1f" % (ph.get("h2d", 0) * 1e3, ph.get("d2h", 0) * 1e3))
    print("   ", " ".join("%s=%.1f" % (a.split(":")[1], b * 1e3) for a, b in tr.items()), " sum %.1f us" % (sum(tr.values()) * 1e3))
