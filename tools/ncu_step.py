#!/usr/bin/env python
"""One pass of every GPU path for an `ncu --set full` capture (developer tool): emit + build of the 10 M-gate headline stream
(multi-kernel pipeline), the fused single-kernel compile of the SHA-256-shaped circuit, and the Kahn levels of the 10 M gate vector."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from c2a_loader import c2a
W = int(sys.argv[1]) if len(sys.argv) > 1 else 18315
ctx = c2a.DeviceContext(0)
wl = c2a.workloads.mimc_chains(W, 91, "late")
k, w, f = c2a.pack_events(np.ascontiguousarray(wl.events), implicit=True)
ins, outs = np.array(sorted(wl.inputs), dtype=np.uint32), np.array(sorted(wl.outputs), dtype=np.uint32)
sha = c2a.workloads.sha256_shaped()
ks, ws, fs = c2a.pack_events(np.ascontiguousarray(sha.events), implicit=True)
si, so = np.array(sorted(sha.inputs), dtype=np.uint32), np.array(sorted(sha.outputs), dtype=np.uint32)
for rep in range(2):                      # pass 0 warms up (slab growth), pass 1 is the one to capture (-s skips pass 0's launches)
    print("PASS", rep, ctx.kernel_launches(), flush=True)
    info = ctx.emit_packed(k, w, f)
    ctx.emitted_build_circuit(ins, outs, want_order=False, want_wires=False)
    gates, _ = ctx.emitted_fetch(want_nodes=False)
    ctx.compile_packed(ks, ws, fs, si, so, want_order=False, want_wires=False)
    ctx.topo_levels(gates, info["node_count"] + 1, level_cap=1 << 20)
print("LAUNCHES", ctx.kernel_launches())
