#!/usr/bin/env python
"""phase timeline of one compile of the headline stream (C2A_PHASE_TIMELINE=1): where does the stream idle? (developer tool, GPU)"""
import os, sys
os.environ["C2A_PHASE_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from c2a_loader import c2a
from circom_2_arithc_b200._lib import PackedEvents, EmitInfo
W = int(sys.argv[1]) if len(sys.argv) > 1 else 18315
ctx = c2a.DeviceContext(0)
wl = c2a.workloads.mimc_chains(W, 91, "late")
k, w, f = c2a.pack_events(np.ascontiguousarray(wl.events), implicit=True)
ins = np.array(sorted(wl.inputs), dtype=np.uint32); outs = np.array(sorted(wl.outputs), dtype=np.uint32)
for i in range(3):
    if i == 2:
        os.environ["C2A_PHASE_TIMELINE"] = "1"
    ctx.compile_packed(k, w, f, ins, outs, want_order=False, want_wires=False, want_gates=False)
