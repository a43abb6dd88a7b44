#!/usr/bin/env python
"""one emit of the 10 M-gate MiMC stream in the given packed form (for ncu captures of k_pk_scatter)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from c2a_loader import c2a
implicit = sys.argv[1] == "packed"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 18315
ctx = c2a.DeviceContext(0)
wl = c2a.workloads.mimc_chains(W, 91, "late")
k, w, f = c2a.pack_events(np.ascontiguousarray(wl.events), implicit=implicit)
for _ in range(3):
    info = ctx.emit_packed(k, w, f)
print(info, ctx.phases().get("k_ev_scatter"), ctx.phases().get("k_ev_scatter_mixed"))
