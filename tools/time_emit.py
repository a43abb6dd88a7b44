#!/usr/bin/env python
"""Phase timings of the device emitter + build on one GPU (diagnostic; bench.py is the contract)."""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from c2a_loader import c2a  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=18315)
    ap.add_argument("--variant", default="late")
    ap.add_argument("--rounds", type=int, default=91)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--levels", action="store_true", help="also time the Kahn level kernel")
    ap.add_argument("--stream", default="packed", choices=["packed", "aos"])
    a = ap.parse_args()
    import torch
    wl = c2a.workloads.mimc_chains(a.chains, rounds=a.rounds, variant=a.variant)
    ev = torch.from_numpy(np.ascontiguousarray(wl.events).view(np.int32)).pin_memory()
    ins = np.array(sorted(wl.inputs), dtype=np.uint32)
    outs = np.array(sorted(wl.outputs), dtype=np.uint32)
    ctx = c2a.DeviceContext(0)
    lib, h, vp = c2a.lib, ctx.handle, C.c_void_p
    from circom_2_arithc_b200._lib import EmitInfo, PackedEvents
    if a.stream == "packed":
        kinds, words, flags = c2a.pack_events(np.ascontiguousarray(wl.events))
        p_k = torch.from_numpy(kinds).pin_memory()
        p_w = torch.from_numpy(words.view(np.int32)).pin_memory()
        pk = PackedEvents(p_k.data_ptr(), p_w.data_ptr(), kinds.shape[0], words.shape[0], flags, 0)
    info = EmitInfo()
    bad = C.c_uint64(0)
    n = ev.shape[0]
    p_order = p_wire = p_new = None
    wc, err = C.c_uint32(0), C.c_uint64(0)
    for rep in range(a.reps):
        t0 = time.perf_counter()
        if a.stream == "packed":
            st = lib.c2a_emit_packed_device(h, C.byref(pk), C.byref(info), C.byref(bad))
        else:
            st = lib.c2a_emit_events_device(h, vp(ev.data_ptr()), n, C.byref(info), C.byref(bad))
        t1 = time.perf_counter()
        assert st == 0, (st, ctx.last_error())
        ph_emit = ctx.phases()
        if p_order is None:
            G, nb = info.n_gates, info.node_count + 1
            p_order = torch.empty(G, dtype=torch.int32).pin_memory()
            p_wire = torch.empty(nb, dtype=torch.int32).pin_memory()
            p_new = torch.empty((G, 4), dtype=torch.int32).pin_memory()
        t2 = time.perf_counter()
        st = lib.c2a_emitted_build_circuit(h, ins.ctypes.data_as(vp), len(ins), outs.ctypes.data_as(vp), len(outs), vp(p_order.data_ptr()),
                                           vp(p_wire.data_ptr()), vp(p_new.data_ptr()), C.byref(wc), C.byref(err))
        t3 = time.perf_counter()
        assert st == 0, (st, ctx.last_error())
        ph_build = ctx.phases()
        print(f"rep {rep}: emit {1e3*(t1-t0):.2f} ms wall, build {1e3*(t3-t2):.2f} ms wall, path={info.path} rounds={info.rounds} G={info.n_gates} "
              f"S={info.signal_bound} C={info.n_connections} eff={info.n_effective} nodes={info.node_count} wires={wc.value}")
        print("   emit phases :", {k: round(v, 3) for k, v in ph_emit.items()})
        print("   build phases:", {k: round(v, 3) for k, v in ph_build.items()})
    if a.levels:
        ctx._emit_info = {"n_gates": info.n_gates, "signal_bound": info.signal_bound}
        gates, _ = ctx.emitted_fetch(want_nodes=False)
        for rep in range(3):
            t0 = time.perf_counter()
            lo, off = ctx.topo_levels(gates, info.node_count + 1)
            print(f"levels: {len(off)-1} levels, {1e3*(time.perf_counter()-t0):.2f} ms wall;", {k: round(v, 3) for k, v in ctx.phases().items()})


if __name__ == "__main__":
    main()
