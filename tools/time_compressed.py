#!/usr/bin/env python
"""where does c2a_emit_compressed_device spend its time? (developer tool, GPU)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from c2a_loader import c2a
ctx = c2a.DeviceContext(0)
src = c2a.workloads.mimc_circom_source(18315, 91)
t0 = time.perf_counter(); dc = c2a.compile(None, source=src, emitter="device", context=ctx); print("walk %.1f ms" % ((time.perf_counter() - t0) * 1e3))
cx = dc.compressed()
print("events", cx.n_events, "words", cx.n_words, "replays", cx.n_replays, "max_gen", cx.max_gen)
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); info = ctx.emit_compressed(cx); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("emit_compressed %.2f ms" % ((t1 - t0) * 1e3), {k: round(v, 3) for k, v in ctx.phases().items() if v > 0.02})
k, w, f = dc._kinds, dc._words, dc._flags
dk, dw = torch.from_numpy(np.ascontiguousarray(k)).cuda(), torch.from_numpy(np.ascontiguousarray(w).view(np.int32)).cuda()
from circom_2_arithc_b200._lib import PackedEvents, EmitInfo
import ctypes as C
pk = PackedEvents(dk.data_ptr(), dw.data_ptr(), len(k), len(w), f, 0)
inf, bad = EmitInfo(), C.c_uint64(0)
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); st = c2a.lib.c2a_emit_packed_resident(ctx.handle, C.byref(pk), C.byref(inf), C.byref(bad)); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("emit_packed_resident of the expanded stream %.2f ms" % ((t1 - t0) * 1e3), st)
