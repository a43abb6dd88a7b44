#!/usr/bin/env python
"""single-kernel vs multi-kernel compile of MiMC-chain streams of growing size: where is the crossover? (developer tool, GPU)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from c2a_loader import c2a
from circom_2_arithc_b200._lib import PackedEvents, EmitInfo, CompileIO
ctx = c2a.DeviceContext(0); h = ctx.handle; lib = c2a.lib
lib.c2a_set_timing(h, 0)
vp = C.c_void_p
for W in [int(x) for x in (sys.argv[1:] or [92, 229, 458, 916, 1832])]:
    wl = c2a.workloads.mimc_chains(W, 91, "late")
    k, w, f = c2a.pack_events(np.ascontiguousarray(wl.events), implicit=True)
    ins = np.array(sorted(wl.inputs), dtype=np.uint32); outs = np.array(sorted(wl.outputs), dtype=np.uint32)
    dk, dw = torch.from_numpy(k).cuda(), torch.from_numpy(w.view(np.int32)).cuda()
    G = wl.n_gates; nb = len(k) - G + 1
    d_order = torch.empty(G, dtype=torch.int32, device="cuda"); d_wire = torch.empty(nb, dtype=torch.int32, device="cuda"); d_new = torch.empty((G, 4), dtype=torch.int32, device="cuda")
    pk = PackedEvents(dk.data_ptr(), dw.data_ptr(), len(k), len(w), f, 0)
    io = CompileIO(ins.ctypes.data_as(vp), outs.ctypes.data_as(vp), len(ins), len(outs), d_order.data_ptr(), d_wire.data_ptr(), d_new.data_ptr(), G, nb, 0)
    info, wc, bad, err = EmitInfo(), C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
    res = {}
    for name, lim in (("fused", 1 << 22), ("multi", 0)):
        lib.c2a_set_fused_limits(lim, 0)
        for _ in range(5):
            st = lib.c2a_compile_packed_resident(h, C.byref(pk), C.byref(io), C.byref(info), C.byref(wc), C.byref(bad), C.byref(err)); assert st == 0
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20):
            lib.c2a_compile_packed_resident(h, C.byref(pk), C.byref(io), C.byref(info), C.byref(wc), C.byref(bad), C.byref(err))
        torch.cuda.synchronize(); res[name] = (time.perf_counter() - t0) / 20 * 1e6
    lib.c2a_set_fused_limits(1 << 22, 0)
    print(f"W={W} gates={G} events={len(k)}: fused {res['fused']:.0f} us   multi-kernel {res['multi']:.0f} us")
