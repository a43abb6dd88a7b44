#!/usr/bin/env python
"""two builds of the shuffled 10 M-gate MiMC vector (BASELINE config 5's stress variant) for ncu captures of k_relax_loop / k_tree_dfs"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from c2a_loader import c2a
W = int(sys.argv[1]) if len(sys.argv) > 1 else 18315
ctx = c2a.DeviceContext(0)
wl = c2a.workloads.mimc_chains(W, 91, "late")
comp = c2a.Compiler(context=ctx)
comp.emit_events(wl.events)
gates = c2a.workloads.shuffle_gates(comp.gate_array(), seed=1)
ins = comp.signal_nodes(np.array(sorted(wl.inputs), dtype=np.uint32))
outs = comp.signal_nodes(np.array(sorted(wl.outputs), dtype=np.uint32))
for _ in range(2):
    order, wire, ng, wc = ctx.build_circuit(gates, comp.node_count + 1, ins, outs)
print("built", len(order), wc, {k: round(v, 3) for k, v in ctx.phases().items() if v > 0.05})
