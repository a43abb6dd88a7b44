#!/usr/bin/env python
"""bench.py — gates/sec of the circom-2-arithc flattening hot path (emit + topo-sort / build_circuit).

    python bench.py --gpus N --steps K --warmup W            # this repo's arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (restated oracle)

Workload (BASELINE.json config 5, the >=1M-gate point the metric is quoted on): synthetic MiMC-chain stream,
W independent chains x 91 rounds x 6 gates emitted in the reference walker's order (circom-2-arithc_b200/
workloads.py).  Default W=18315 -> 10.0 M gates; 'late' variant = component inputs wired after the body, so the
DFS post-order is NOT the identity and the full sort path runs.

A step = one pass of the hot path (emit + build_circuit) over one circuit:
  value  event stream resident in HBM -> c2a_emit_events_resident (device emitter) -> c2a_emitted_build_circuit_device
         (K1..K7); results stay in HBM; CUDA events on the handle's stream
  e2e    event stream in PINNED HOST memory -> c2a_emit_events_device -> c2a_emitted_build_circuit into pinned host
         buffers (H2D of the events and D2H of order / wire map / new gates inside the timed region)
  e2e_host_emitter  (one step, for comparison) the same circuit through the host union-find emitter + c2a_build_circuit
N>1: every rank owns one independent component subtree (its own W chains; weak scaling), ranks exchange their
input/intermediate/output wire counts with one NCCL all-gather and rebase their wires to the global numbering.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic bytes of each kernel per launch (SURVEY.md §8d accounting: every array element counted once
# per required read and once per required write, atomics = read + write of the word).  c = the workload's counts.
def alg_bytes(kernel, c):
    G, NB, n, S, C, Ceff, ns = c["G"], c["NB"], c["n"], c["S"], c["C"], c["Ceff"], c["n_sig"]
    order = 0 if c["identity"] else 4 * G            # sorted passes also read order[]
    table = {
        # ---- device emitter (c2a_emit.cuh)
        # count pass, then scatter to sig_t + sig_meta | egates + gate_t | conn + conn_t + conn_sb
        "emit:k_ev_count": c["stream_bytes_count"] + 8 * ((n + 1023) // 1024),   # AoS: 16 B/event; packed: the kind bytes only
        # dense packed stream: validated in place (no event-time arrays, no declaration table, no E2 kernels)
        "emit:k_ev_scatter": c["stream_bytes"] + 16 * ((n + 1023) // 1024) + (ns * 4 + G * (16 + 1) + C * (8 + 4) if c["dense"] else
                                                                             ns * (4 + 8) + G * (16 + 4) + C * (8 + 4 + 4)),
        "emit:k_ev_check_gates": 0 if c["dense"] else G * (16 + 4 + 12 + 1),
        "emit:k_ev_check_conns": 0 if c["dense"] else C * (8 + 4 + 8),
        "emit:k_msf_pick": C * (8 + 16),                          # first round: conn, 2 RED.MIN best (later rounds run on the shrunken list)
        "emit:k_msf_hook": C * (8 + 8 + 4 + 4),                   # first round: conn, 2 best, parent, eff
        "emit:k_scan_u32": 8 * (C // 32 + 1) + 16 * ((n + 1023) // 1024),   # effective-connection bitmap ranks + the two tile-count scans
        "emit:k_ev_nid_edges": C * (4 + 8) + Ceff * (4 + 8),      # conn_sb, conn (+ bitmap words, L2) ; parent chase + atomicMax on the root
        "emit:k_ev_finalize": S * ((4 if c["dense"] else 4 + 8) + 1 + 4 + 4 + 4) + (G + c["n_const"]) * 8,   # sig_t + meta (dense ids: one 4-byte record), outmark, parent, id word, nos + screen atomics
        "emit:k_ev_gates": G * (16 + 12 + 16 + 4),                # signal-id gate, 3 node gathers, node-id gate, RED.MAX producer[out] (K1 of the build)
        "emit:init": (0 if c["dense"] else 4 * (n + (1 << 20))) + S * (1 + 4 + 4 + 4) + 4 * NB,   # memsets: sig_t (bound-sized), outmark, best, parent iota, {nid,cnt}, eff; producer[] (side stream)
        # ---- build_circuit (c2a_device.cu)
        "k_producer": G * (16 + 4),                               # read gate, RED.MAX producer[out] (only for gates that did not come from the device emitter)
        "k_deps": G * (16 + 8 + 8),                               # read gate, 2 producer gathers, write dep pair (+ the forward-edge list)
        "k_relax": 0,                                             # data-driven from the forward-edge list k_deps collects: traffic ~ the moved cones only
        "k_sizes": G * (4 + 4),
        "k_scan_u32": G * (4 + 4) + (0 if "W" not in c else 8 * c["W"]),   # block-offset scan + bitmap rank scan
        "k_roots": G * (4 + 8 + 4),
        "k_tree_dfs": 0,
        "k_wire_first": G * (16 + 12) + order,                    # read gate, 3 RED.MIN on wire[]
        "k_wire_mark": 4 * NB + 8 * c["W"],                       # stream wire[], set first-appearance bits (bitmap RMW)
        "k_wire_assign": 4 * NB + 4 * c["n_mid"] + 8 * c["W"],    # stream wire[], bitmap + rank-prefix lookups, one write per numbered node
        "k_gather": G * (16 + 12 + 16) + order,                   # read gate, 3 wire gathers, write new gate
        "init": 4 * NB + (0 if c["identity"] else 9 * G + G // 8),   # fill wire[] (side stream); sort scratch (r, size_off, state, inq)
    }
    return table.get(kernel, 0)


def clocks_sampler(stop, out, index):
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return
    def rd():
        for line in p.stdout:
            out.append(line.strip())
    t = threading.Thread(target=rd, daemon=True)
    t.start()
    stop.wait()
    p.terminate()


def summarize_clocks(lines):
    sm, mx, reasons = [], 0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        f = [x.strip() for x in ln.split(",")]
        if len(f) < 7:
            continue
        try:
            sm.append(float(f[0]))
            mx = max(mx, float(f[1]))
        except ValueError:
            continue
        for n, v in zip(names, f[3:7]):
            if v.lower().startswith("active"):
                reasons.add(n)
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(torch, local_rank):
    """Best effort: run this rank (and first-touch its pinned buffers) on the CPUs next to its GPU.  With 4-8 ranks the e2e leg moves
    ~0.4 GB per rank and step over PCIe; buffers on the wrong socket make every copy cross the inter-socket link."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        dev = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        base = f"/sys/bus/pci/devices/{dev}"
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        if node >= 0 and cpus:
            os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or cpus)
            return node
        return "sysfs numa_node = %d for %s (a single-NUMA-node VM exposes no GPU affinity): nothing to bind" % (node, dev)
    except Exception as e:  # noqa: BLE001
        return "unavailable (%s)" % type(e).__name__


def cpu_reference_measure(workloads, wl_full, sample_chains, variant, rounds, backend_gates=None):
    """The reference's CPU path, restated (oracle, faithful data structures), single thread like the reference.
    emit: O(G*S) scans (src/compiler.rs:185-195, 219-226, 260-270) on a bounded prefix of the workload;
    back end: HashMap producer map + DFS + first-seen numbering + gather (src/compiler.rs:388-464)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib as orc
    wl = workloads.mimc_chains(sample_chains, rounds=rounds, variant=variant)
    oc = orc.OracleCompiler()
    t0 = time.perf_counter()
    oc.emit_events(wl.events)
    t_emit = time.perf_counter() - t0
    gates = oc.gate_array()
    ins = np.array([oc.signal_node(s) for s in sorted(wl.inputs)], dtype=np.uint32)
    outs = np.array([oc.signal_node(s) for s in sorted(wl.outputs)], dtype=np.uint32)
    t_back, st = orc.backend_time(gates, ins, outs, reps=3)
    assert st == 0
    res = {"sample_gates": int(gates.shape[0]), "emit_s": t_emit, "backend_s": t_back,
           "gates_per_s": gates.shape[0] / (t_emit + t_back)}
    if backend_gates is not None:
        g, i, o = backend_gates
        tb, st = orc.backend_time(g, i, o, reps=1)
        assert st == 0
        res["backend_only_full_gates"] = int(g.shape[0])
        res["backend_only_gates_per_s"] = g.shape[0] / tb
    return res


def load_workloads_without_native():
    """workloads.py (numpy only) loaded under a stub package, so that the reference arm's process never maps libc2a.so
    (importing the real package loads it)."""
    import importlib.util
    import types
    pkg_dir = os.path.join(ROOT, "circom-2-arithc_b200")
    stub = types.ModuleType("c2a_workloads_only")
    stub.__path__ = [pkg_dir]
    sys.modules["c2a_workloads_only"] = stub
    mods = {}
    for name in ("gate_types", "workloads"):
        spec = importlib.util.spec_from_file_location(f"c2a_workloads_only.{name}", os.path.join(pkg_dir, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["workloads"]


# algorithmic bytes per gate of the phases.  The exact-order sort this build runs (K2 + K5a-c) moves
# deps 32 + scratch init 9 + sizes 8 + offsets scan 8 + roots 16 = 73 B/gate; Kahn K1-K4 (CSR) 104 B/gate.
ALG_SORT_BYTES_PER_GATE = 73
ALG_KAHN_BYTES_PER_GATE = 104
SORT_PHASES = ("k_deps", "k_relax", "k_sizes", "k_roots", "k_tree_dfs")


class StagedCircuit:
    """One workload staged for the two measured forms: the packed stream in pinned host memory (e2e) and resident in HBM (value),
    plus pinned / device result buffers sized by one untimed pass."""

    def __init__(self, c2a, torch, ctx, dev, wl):
        import numpy as np
        from circom_2_arithc_b200._lib import EmitInfo, PackedEvents
        self.c2a, self.torch, self.ctx, self.lib, self.h, self.wl = c2a, torch, ctx, c2a.lib, ctx.handle, wl
        ev = np.ascontiguousarray(wl.events)
        self.n_ev = int(ev.shape[0])
        kinds, words, flags = c2a.pack_events(ev, implicit=True)
        self.p_kinds = torch.from_numpy(kinds).pin_memory()
        self.p_words = torch.from_numpy(words.view(np.int32)).pin_memory()
        self.d_kinds, self.d_words = self.p_kinds.to(dev), self.p_words.to(dev)
        nw = int(words.shape[0])
        self.pk_host = PackedEvents(self.p_kinds.data_ptr(), self.p_words.data_ptr(), self.n_ev, nw, flags, 0)
        self.pk_dev = PackedEvents(self.d_kinds.data_ptr(), self.d_words.data_ptr(), self.n_ev, nw, flags, 0)
        self.stream_bytes = self.n_ev + 4 * nw
        self.ins = np.array(sorted(wl.inputs), dtype=np.uint32)
        self.outs = np.array(sorted(wl.outputs), dtype=np.uint32)
        kk = ev[:, 0] & 0xFF
        self.named = np.concatenate([self.ins, self.outs, ev[kk == 1, 1]]).astype(np.uint32)
        self.info, self.bad, self.wc, self.err = EmitInfo(), C.c_uint64(0), C.c_uint32(0), C.c_uint64(0)
        st = self.lib.c2a_emit_packed_resident(self.h, C.byref(self.pk_dev), C.byref(self.info), C.byref(self.bad))
        if st != 0 or self.info.path != 1:
            raise RuntimeError(f"{wl.name}: emit -> {st} path {self.info.path}: {ctx.last_error()}")
        self.G, self.nb = int(self.info.n_gates), int(self.info.node_count) + 1
        G, nb = self.G, self.nb
        self.d_order = torch.empty(max(G, 1), dtype=torch.int32, device=dev)
        self.d_wire = torch.empty(nb, dtype=torch.int32, device=dev)
        self.d_new = torch.empty((max(G, 1), 4), dtype=torch.int32, device=dev)
        self.p_new = torch.empty((max(G, 1), 4), dtype=torch.int32).pin_memory()
        self.p_named = torch.from_numpy(self.named.view(np.int32)).pin_memory()
        self.p_named_w = torch.empty(max(len(self.named), 1), dtype=torch.int32).pin_memory()

    def step_resident(self):
        """emit + build, packed stream resident in HBM, results stay in HBM"""
        lib, h, vp = self.lib, self.h, C.c_void_p
        st = lib.c2a_emit_packed_resident(h, C.byref(self.pk_dev), C.byref(self.info), C.byref(self.bad))
        if st == 0:
            st = lib.c2a_emitted_build_circuit_device(h, self.ins.ctypes.data_as(vp), len(self.ins), self.outs.ctypes.data_as(vp), len(self.outs),
                                                      vp(self.d_order.data_ptr()), vp(self.d_wire.data_ptr()), vp(self.d_new.data_ptr()),
                                                      C.byref(self.wc), C.byref(self.err))
        if st != 0:
            raise RuntimeError(f"{self.wl.name}: resident step -> {st}: {self.ctx.last_error()}")

    def _io(self, order, wire, new, gcap, wcap):
        from circom_2_arithc_b200._lib import CompileIO
        vp = C.c_void_p
        return CompileIO(self.ins.ctypes.data_as(vp), self.outs.ctypes.data_as(vp), len(self.ins), len(self.outs), order, wire, new, gcap, wcap, 0)

    def step_compile_resident(self):
        """c2a_compile_packed_resident: emit + build in ONE call (one synchronisation; the single cooperative kernel for circuits
        up to ~1 M gates), packed stream resident in HBM, results stay in HBM"""
        if not hasattr(self, "_io_res"):
            self._io_res = self._io(self.d_order.data_ptr(), self.d_wire.data_ptr(), self.d_new.data_ptr(), self.G, self.nb)
        st = self.lib.c2a_compile_packed_resident(self.h, C.byref(self.pk_dev), C.byref(self._io_res), C.byref(self.info), C.byref(self.wc), C.byref(self.bad), C.byref(self.err))
        if st != 0:
            raise RuntimeError(f"{self.wl.name}: c2a_compile_packed_resident -> {st}: {self.ctx.last_error()}")

    def step_compile_e2e(self):
        """c2a_compile_packed: the same from / into pinned host buffers (H2D of the stream, D2H of the renumbered gates inside), then the
        named wires"""
        vp = C.c_void_p
        if not hasattr(self, "_io_host"):
            self._io_host = self._io(None, None, self.p_new.data_ptr(), self.G, 0)
        st = self.lib.c2a_compile_packed(self.h, C.byref(self.pk_host), C.byref(self._io_host), C.byref(self.info), C.byref(self.wc), C.byref(self.bad), C.byref(self.err))
        if st == 0:
            st = self.lib.c2a_emitted_signal_wires(self.h, vp(self.p_named.data_ptr()), len(self.named), vp(self.p_named_w.data_ptr()))
        if st != 0:
            raise RuntimeError(f"{self.wl.name}: c2a_compile_packed -> {st}: {self.ctx.last_error()}")

    def step_e2e(self):
        """the same through host buffers: packed stream from pinned memory, renumbered gates + named wires into pinned memory"""
        lib, h, vp = self.lib, self.h, C.c_void_p
        st = lib.c2a_emit_packed_device(h, C.byref(self.pk_host), C.byref(self.info), C.byref(self.bad))
        if st == 0:
            st = lib.c2a_emitted_build_circuit(h, self.ins.ctypes.data_as(vp), len(self.ins), self.outs.ctypes.data_as(vp), len(self.outs), None, None,
                                               vp(self.p_new.data_ptr()), C.byref(self.wc), C.byref(self.err))
        if st == 0:
            st = lib.c2a_emitted_signal_wires(h, vp(self.p_named.data_ptr()), len(self.named), vp(self.p_named_w.data_ptr()))
        if st != 0:
            raise RuntimeError(f"{self.wl.name}: e2e step -> {st}: {self.ctx.last_error()}")

    def e2e_bytes(self):
        return self.stream_bytes + 4 * (len(self.ins) + len(self.outs)) + 4 * len(self.named), 16 * self.G + 4 * len(self.named) + 4 * 32 * 3


def time_on_stream(torch, stream, fn, steps, warmup, flush=None):
    """CUDA events on `stream` around `steps` calls of fn (after `warmup` untimed ones).  flush: a callable run between the timed
    calls OUTSIDE the event pairs (evicts the L2); then every call gets its own event pair and the mean is returned."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    tot = 0.0
    for _ in range(steps):
        flush()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


def circuit_leg(c2a, torch, ctx, dev, stream, wl, steps, peak, flush, cpu_backend=True, orc=None, concurrent=0):
    """One BASELINE config as a sub-record: value (resident), value with the L2 flushed between steps, e2e (host buffers), the
    exact-order sort's achieved HBM GB/s, and the CPU oracle's back end on the very same gate vector."""
    import numpy as np
    lib, h = c2a.lib, ctx.handle
    sc = StagedCircuit(c2a, torch, ctx, dev, wl)
    lib.c2a_set_timing(h, 0)
    # the product's call for a whole recording: c2a_compile_packed* (one call, one synchronisation)
    ms = time_on_stream(torch, stream, sc.step_compile_resident, steps, 3)
    ms_cold = time_on_stream(torch, stream, sc.step_compile_resident, max(3, steps // 4), 1, flush=flush)
    sc.step_compile_e2e()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        sc.step_compile_e2e()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / steps
    # throughput with several circuits in flight: T handles on T host threads, each compiling its own copy of this circuit (the
    # single-kernel path occupies a few CTAs per small circuit, so independent compilations share the GPU)
    conc = None
    if concurrent and sc.G <= 500_000:
        T = int(concurrent)
        ctxs = [c2a.DeviceContext(ctx.device) for _ in range(T)]
        scs = [StagedCircuit(c2a, torch, cx_, dev, wl) for cx_ in ctxs]
        for cx_ in ctxs:
            lib.c2a_set_timing(cx_.handle, 0)
        reps = max(20, steps)

        def worker(s_):
            for _ in range(reps):
                s_.step_compile_resident()

        for s_ in scs:
            s_.step_compile_resident()
        torch.cuda.synchronize()
        ths = [threading.Thread(target=worker, args=(s_,)) for s_ in scs]
        t0 = time.perf_counter()
        for t_ in ths:
            t_.start()
        for t_ in ths:
            t_.join()
        torch.cuda.synchronize()
        dtc = time.perf_counter() - t0
        conc = {"value": T * reps * sc.G / dtc, "unit": "gates/s", "handles_in_flight": T, "circuits": T * reps, "wall_s": dtc,
                "note": "T handles on T host threads, same call as `value` (c2a_compile_packed_resident), wall clock around all of them"}
        del scs, ctxs
    # for comparison: the two-call multi-kernel pipeline on the same circuit (what round 1 measured)
    ms_multi = time_on_stream(torch, stream, sc.step_resident, max(3, steps // 2), 2)
    lib.c2a_set_timing(h, 1)
    lib.c2a_set_timing_only(h, None)
    l0 = ctx.kernel_launches()
    sc.step_compile_resident()
    launches = ctx.kernel_launches() - l0
    fused = "k_fused_compile" in ctx.phases()
    fused_ms = ctx.phases().get("k_fused_compile")
    sc.step_resident()
    ph = ctx.phases()
    sort_ms = sum(ph.get(k, 0.0) for k in SORT_PHASES)
    h2d, d2h = sc.e2e_bytes()
    rec = {"workload": wl.name, "gates": sc.G, "events": sc.n_ev, "node_bound": sc.nb, "value": sc.G / (ms * 1e-3), "unit": "gates/s", "ms_per_step": ms,
           "value_l2_flushed": sc.G / (ms_cold * 1e-3), "ms_per_step_l2_flushed": ms_cold, "steps": steps, "gpu_launches_per_step": int(launches),
           "path": "c2a_compile_packed_resident: " + ("one cooperative kernel (csrc/c2a_fused.cuh)" if fused else "multi-kernel pipeline, one call"),
           "fused_kernel_ms": fused_ms, "multi_kernel_two_call_ms_per_step": ms_multi,
           **({"value_concurrent": conc} if conc else {}),
           "e2e": {"value": sc.G / e2e_s, "unit": "gates/s", "s_per_step": e2e_s, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
           "topo_sort_ms": sort_ms, "topo_hbm_gbs": (ALG_SORT_BYTES_PER_GATE * sc.G / (sort_ms * 1e-3) / 1e9) if sort_ms > 0 else None,
           "topo_frac_of_peak": (ALG_SORT_BYTES_PER_GATE * sc.G / (sort_ms * 1e-3) / 1e9 / peak) if sort_ms > 0 else None,
           "l2": "circuit smaller than the 126 MB L2: `value` is L2-warm back-to-back steps, `value_l2_flushed` evicts the L2 (256 MB write) before every step"
                 if sc.stream_bytes + 16 * sc.G < 100e6 else "inputs larger than L2; no flush needed"}
    if cpu_backend and orc is not None:
        sc.ctx._emit_info = {"n_gates": sc.G, "signal_bound": int(sc.info.signal_bound)}
        gates, nos = ctx.emitted_fetch()
        tb, st = orc.backend_time(gates, nos[sc.ins], nos[sc.outs], reps=3 if sc.G < 2_000_000 else 1)
        assert st == 0
        rec["cpu_backend_gates_per_s"] = sc.G / tb
        rec["cpu_backend_note"] = "oracle back end (HashMap producer map + DFS + first-seen numbering + gather), 1 thread, same gate vector"
        # parity on the spot: the GPU result of this very circuit against the oracle
        st_, _e, o_order, _w, o_gates, o_wc = orc.backend_raw(gates, sc.nb, nos[sc.ins], nos[sc.outs])
        ok = (st_ == 0 and np.array_equal(sc.d_order[:sc.G].cpu().numpy().astype(np.uint32), o_order)
              and np.array_equal(sc.d_new[:sc.G].cpu().numpy().astype(np.uint32), o_gates) and o_wc == sc.wc.value)
        rec["parity_vs_oracle"] = "ok" if ok else "MISMATCH"
        assert ok, f"{wl.name}: GPU result differs from the oracle"
    return rec, sc


def backend_leg(c2a, torch, ctx, dev, stream, name, gates, nb, ins_nodes, outs_nodes, steps, peak, orc=None, oracle_reps=1):
    """c2a_build_circuit_device on a caller's gate array resident in HBM (K1..K7): the back end alone."""
    import numpy as np
    lib, h, vp = c2a.lib, ctx.handle, C.c_void_p
    G = int(gates.shape[0])
    d_gates = torch.from_numpy(np.ascontiguousarray(gates).view(np.int32)).to(dev)
    d_order = torch.empty(G, dtype=torch.int32, device=dev)
    d_wire = torch.empty(nb, dtype=torch.int32, device=dev)
    d_new = torch.empty((G, 4), dtype=torch.int32, device=dev)
    wc, err = C.c_uint32(0), C.c_uint64(0)
    ins_nodes, outs_nodes = np.ascontiguousarray(ins_nodes, dtype=np.uint32), np.ascontiguousarray(outs_nodes, dtype=np.uint32)

    def step():
        st = lib.c2a_build_circuit_device(h, vp(d_gates.data_ptr()), G, nb, ins_nodes.ctypes.data_as(vp), len(ins_nodes), outs_nodes.ctypes.data_as(vp),
                                          len(outs_nodes), vp(d_order.data_ptr()), vp(d_wire.data_ptr()), vp(d_new.data_ptr()), C.byref(wc), C.byref(err))
        if st != 0:
            raise RuntimeError(f"{name}: c2a_build_circuit_device -> {st}: {ctx.last_error()}")

    lib.c2a_set_timing(h, 0)
    ms = time_on_stream(torch, stream, step, steps, 2)
    lib.c2a_set_timing(h, 1)
    lib.c2a_set_timing_only(h, None)
    step()
    ph = ctx.phases()
    sort_ms = sum(ph.get(k, 0.0) for k in SORT_PHASES) + ph.get("k_producer", 0.0)
    rec = {"workload": name, "gates": G, "value": G / (ms * 1e-3), "unit": "gates/s", "ms_per_step": ms, "steps": steps,
           "scope": "gate vector resident in HBM -> c2a_build_circuit_device (producer map, deps, exact DFS order, wire numbering, gather); results stay in HBM",
           "relax_fallback_rounds": int(ph.get("n_relax_fallback_rounds", 0)),
           "topo_sort_ms": sort_ms, "topo_hbm_gbs": ((ALG_SORT_BYTES_PER_GATE + 20) * G / (sort_ms * 1e-3) / 1e9) if sort_ms > 0 else None,
           "phases_ms": {k: round(v, 4) for k, v in ph.items()}}
    if rec["topo_hbm_gbs"]:
        rec["topo_frac_of_peak"] = rec["topo_hbm_gbs"] / peak
    if orc is not None:
        tb, st = orc.backend_time(gates, ins_nodes, outs_nodes, reps=oracle_reps)
        assert st == 0
        rec["cpu_backend_gates_per_s"] = G / tb
        rec["speedup_vs_cpu_backend"] = rec["value"] / rec["cpu_backend_gates_per_s"]
    return rec, (d_order, d_new, wc)


def kahn_leg(c2a, torch, ctx, dev, name, gates, nb, steps, peak):
    """K1..K4 (producer map, deps, consumer CSR, asynchronous Kahn walk, counting sort by level) on a resident gate array:
    SURVEY 8d's 104 B/gate against the measured copy peak."""
    import numpy as np
    lib, h, vp = c2a.lib, ctx.handle, C.c_void_p
    G = int(gates.shape[0])
    d_gates = torch.from_numpy(np.ascontiguousarray(gates).view(np.int32)).to(dev)
    d_lo = torch.empty(G, dtype=torch.int32, device=dev)
    cap = 1 << 20
    d_off = torch.empty(cap + 2, dtype=torch.int32, device=dev)
    nl, err = C.c_uint32(0), C.c_uint64(0)
    best_call, best_kern, phases = 1e30, 1e30, {}
    lib.c2a_set_timing(h, 1)
    lib.c2a_set_timing_only(h, None)
    for i in range(steps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = lib.c2a_topo_levels_device(h, vp(d_gates.data_ptr()), G, nb, vp(d_lo.data_ptr()), vp(d_off.data_ptr()), cap, C.byref(nl), C.byref(err))
        dt = (time.perf_counter() - t0) * 1e3
        if st != 0:
            raise RuntimeError(f"{name}: c2a_topo_levels_device -> {st}: {ctx.last_error()}")
        ph = ctx.phases()
        kern = sum(v for k, v in ph.items() if k.startswith("k_") or k == "init")
        if i and kern < best_kern:
            best_kern, phases = kern, ph
        if i:
            best_call = min(best_call, dt)
    ab = ALG_KAHN_BYTES_PER_GATE * G
    return {"workload": name, "gates": G, "levels": int(nl.value), "ms_kernels": best_kern, "ms_call": best_call, "alg_bytes": ab,
            "achieved_gbs": ab / (best_kern * 1e-3) / 1e9, "frac": ab / (best_kern * 1e-3) / 1e9 / peak,
            "latency_floor_note": "%d levels: a chain of %d dependent hops bounds the walk from below (~1 us per DRAM hop)" % (nl.value, nl.value),
            "phases_ms": {k: round(v, 4) for k, v in phases.items()}}



def multi_gpu_parity(c2a, torch, dist, ctx, dev, stream, rank, world, variant, rounds):
    """Driver-visible N>1 parity, run before the timed region on every rank:
      (a) sharding.build_circuit_sharded of ONE circuit (plan_shards cuts, per-shard build, one all-gather of counts, rebase) against the
          single-rank build of the same gate vector and against the oracle;
      (b) the bench's own weak-scaling path (every rank emits and builds its own component subtree, all-gather of counts, offsets applied
          inside the gather kernel) against the oracle's build of the concatenated circuit (inputs of all ranks, intermediates, outputs).
    Returns "ok" or raises."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc
    lib, h, vp = c2a.lib, ctx.handle, C.c_void_p
    ok = True
    # ---- (a) one circuit, sharded
    cases = [c2a.workloads.mimc_chains(50 * world, rounds=rounds, variant=variant)]
    if world == 2:
        cases.append(c2a.workloads.keccak_shaped(instances=2, rounds=3))
    for wl in cases:
        comp = c2a.Compiler(context=ctx)
        comp.emit_events(wl.events)
        g, nb = comp.gate_array(), comp.node_count + 1
        ins = comp.signal_nodes(np.array(sorted(wl.inputs), dtype=np.uint32))
        outs = comp.signal_nodes(np.array(sorted(wl.outputs), dtype=np.uint32))
        order, wire, ng, wc, plan = c2a.sharding.build_circuit_sharded(g, nb, ins, outs, ctx=ctx)
        o1, w1, g1, wc1 = ctx.build_circuit(g, nb, ins, outs)
        st, _, o_order, o_wire, o_gates, o_wc = orc.backend_raw(g, nb, ins, outs)
        ok = ok and st == 0 and len(plan) == world and wc == wc1 == o_wc
        ok = ok and np.array_equal(order, o1) and np.array_equal(order, o_order) and np.array_equal(ng, g1) and np.array_equal(ng, o_gates)
        ok = ok and np.array_equal(wire, w1) and np.array_equal(wire, o_wire)
        del comp
    # ---- (b) the weak-scaling path of this bench on a small workload
    sc = StagedCircuit(c2a, torch, ctx, dev, c2a.workloads.mimc_chains(40 + rank, rounds=rounds, variant=variant))   # ragged: a different size per rank
    st = lib.c2a_emit_packed_resident(h, C.byref(sc.pk_dev), C.byref(sc.info), C.byref(sc.bad))
    assert st == 0 and sc.info.path == 1
    sc.ctx._emit_info = {"n_gates": sc.G, "signal_bound": int(sc.info.signal_bound)}
    gates_local, nos = ctx.emitted_fetch()
    st = lib.c2a_emitted_build_circuit_device(h, sc.ins.ctypes.data_as(vp), len(sc.ins), sc.outs.ctypes.data_as(vp), len(sc.outs), vp(sc.d_order.data_ptr()),
                                              vp(sc.d_wire.data_ptr()), None, C.byref(sc.wc), C.byref(sc.err))
    assert st == 0, ctx.last_error()
    h_counts = torch.tensor([len(sc.ins), sc.wc.value - len(sc.ins) - len(sc.outs), len(sc.outs), sc.G], dtype=torch.int64).pin_memory()
    d_counts = torch.zeros(4, dtype=torch.int64, device=dev)
    d_all = torch.zeros(4 * world, dtype=torch.int64, device=dev)
    with torch.cuda.stream(stream):
        d_counts.copy_(h_counts, non_blocking=True)
        dist.all_gather_into_tensor(d_all, d_counts)
    st = lib.c2a_emitted_gather_device(h, vp(sc.d_order.data_ptr()), vp(sc.d_new.data_ptr()), vp(d_all.data_ptr()), rank, world)
    assert st == 0, ctx.last_error()
    stream.synchronize()
    mine = {"gates": gates_local, "nb": sc.nb, "ins": nos[sc.ins], "outs": nos[sc.outs], "new": sc.d_new[:sc.G].cpu().numpy().astype(np.uint32),
            "order": sc.d_order[:sc.G].cpu().numpy().astype(np.uint32)}
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    if rank == 0:
        # concatenated circuit: rank r's node ids shifted by the node bounds before it; inputs / outputs listed rank by rank
        node_base = np.cumsum([0] + [p_["nb"] for p_ in parts])
        gg = np.concatenate([p_["gates"] + np.array([0, b, b, b], dtype=np.uint32) for p_, b in zip(parts, node_base[:-1])])
        gi = np.concatenate([p_["ins"] + np.uint32(b) for p_, b in zip(parts, node_base[:-1])])
        go = np.concatenate([p_["outs"] + np.uint32(b) for p_, b in zip(parts, node_base[:-1])])
        st, _, o_order, _w, o_gates, _wc = orc.backend_raw(gg, int(node_base[-1]), gi, go)
        ok = ok and st == 0 and np.array_equal(np.concatenate([p_["order"] for p_ in parts]), o_order)
        ok = ok and np.array_equal(np.concatenate([p_["new"] for p_ in parts]), o_gates)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) != 1:
        raise RuntimeError("multi-GPU parity check failed: the stitched N-rank result differs from the single-rank / oracle result")
    return "ok"



def strong_scaling_leg(c2a, torch, dist, ctx, dev, stream, rank, world, wl, steps):
    """ONE circuit (the headline workload, every rank holds the whole packed stream) built by `world` GPUs: every rank replays the
    stream (the emit is replicated: node ids are global prefix counts), plans the shards ON THE DEVICE (c2a_plan_shards_device),
    builds its own gate range (c2a_emitted_build_range_device, shared I/O lists), the ranks all-gather their (n_in, n_mid, n_out, G)
    and shift their wire ids / gate indices to the global numbering.  Device time, max over ranks; the stitched result is checked
    against a single-rank build once before the timed steps."""
    import numpy as np
    lib, h, vp = c2a.lib, ctx.handle, C.c_void_p
    sc = StagedCircuit(c2a, torch, ctx, dev, wl)
    G, nb = sc.G, sc.nb
    ins_n = None
    d_order = torch.empty(G, dtype=torch.int32, device=dev)
    d_wire = torch.empty(nb, dtype=torch.int32, device=dev)
    d_new = torch.empty((G, 4), dtype=torch.int32, device=dev)
    d_counts = torch.zeros(4, dtype=torch.int64, device=dev)
    d_all = torch.zeros(4 * world, dtype=torch.int64, device=dev)
    h_counts = torch.zeros(4, dtype=torch.int64).pin_memory()
    wc, err = C.c_uint32(0), C.c_uint64(0)
    bounds = (C.c_uint64 * (world + 1))()
    nsh = C.c_uint32(0)
    t_phase = {"emit": 0.0, "plan": 0.0, "build": 0.0, "reconcile": 0.0}
    state = {}

    def step(timed=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if timed else None
        if timed:
            ev[0].record(stream)
        st = lib.c2a_emit_packed_resident(h, C.byref(sc.pk_dev), C.byref(sc.info), C.byref(sc.bad))
        assert st == 0 and sc.info.path == 1, ctx.last_error()
        if timed:
            ev[1].record(stream)
        if ins_n is None:
            return
        gp = lib.c2a_emitted_gates_device(h)
        st = lib.c2a_plan_shards_device(h, vp(gp), G, nb, ins_n.ctypes.data_as(vp), len(ins_n), outs_n.ctypes.data_as(vp), len(outs_n), world, bounds, C.byref(nsh))
        assert st == 0 and nsh.value == world, (st, nsh.value, ctx.last_error())
        if timed:
            ev[2].record(stream)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        st = lib.c2a_emitted_build_range_device(h, lo, hi, sc.ins.ctypes.data_as(vp), len(sc.ins), sc.outs.ctypes.data_as(vp), len(sc.outs),
                                                vp(d_order.data_ptr()), vp(d_wire.data_ptr()), vp(d_new.data_ptr()), C.byref(wc), C.byref(err))
        assert st == 0, ctx.last_error()
        if timed:
            ev[3].record(stream)
        n_mid = wc.value - len(sc.ins) - len(sc.outs)
        h_counts[0], h_counts[1], h_counts[2], h_counts[3] = len(sc.ins), n_mid, len(sc.outs), hi - lo
        with torch.cuda.stream(stream):
            d_counts.copy_(h_counts, non_blocking=True)
            dist.all_gather_into_tensor(d_all, d_counts)
        stream.synchronize()
        counts = d_all.view(world, 4).cpu().numpy()
        off_in, off_mid, off_out, gate_base = c2a.sharding.rebase_offsets(counts, rank, shared_io=True)
        st = lib.c2a_rebase_wires_device(h, vp(d_new.data_ptr()), vp(d_order.data_ptr()), hi - lo, len(sc.ins), n_mid, off_in, off_mid, off_out, gate_base)
        assert st == 0, ctx.last_error()
        if timed:
            ev[4].record(stream)
            torch.cuda.synchronize()
            for k_, (a, b) in zip(("emit", "plan", "build", "reconcile"), zip(ev[:-1], ev[1:])):
                t_phase[k_] += a.elapsed_time(b)
        state.update(lo=lo, hi=hi, n_mid=n_mid, counts=counts)

    step()      # emit only: the I/O node ids of the plan come from the resident signal -> node map
    sc.ctx._emit_info = {"n_gates": G, "signal_bound": int(sc.info.signal_bound)}
    nos = ctx.emitted_fetch(want_gates=False)[1]
    ins_n, outs_n = np.ascontiguousarray(nos[sc.ins]), np.ascontiguousarray(nos[sc.outs])
    # ---- parity of the stitched result against the single-rank build of the same circuit (order + renumbered gates, by slice)
    step()
    lo, hi = state["lo"], state["hi"]
    mine_o = d_order[:hi - lo].cpu().numpy().astype(np.uint32)
    mine_g = d_new[:hi - lo].cpu().numpy().astype(np.uint32)
    st = lib.c2a_emit_packed_resident(h, C.byref(sc.pk_dev), C.byref(sc.info), C.byref(sc.bad))
    st = st or lib.c2a_emitted_build_circuit_device(h, sc.ins.ctypes.data_as(vp), len(sc.ins), sc.outs.ctypes.data_as(vp), len(sc.outs), vp(sc.d_order.data_ptr()),
                                                    vp(sc.d_wire.data_ptr()), vp(sc.d_new.data_ptr()), C.byref(wc), C.byref(err))
    assert st == 0, ctx.last_error()
    ok = np.array_equal(mine_o, sc.d_order[lo:hi].cpu().numpy().astype(np.uint32)) and np.array_equal(mine_g, sc.d_new[lo:hi].cpu().numpy().astype(np.uint32))
    ok = ok and wc.value == len(sc.ins) + int(state["counts"][:, 1].sum()) + len(sc.outs)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) != 1:
        raise RuntimeError("strong scaling: the stitched sharded build differs from the single-rank build")
    lib.c2a_set_timing(h, 0)
    for _ in range(2):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    for k_ in t_phase:
        t_phase[k_] = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    for _ in range(2):
        step(timed=True)
    lib.c2a_set_timing(h, 1)
    ms = float(t.item())
    return {"workload": wl.name, "gates_total": G, "n_gpus": world, "value": G / (ms * 1e-3), "unit": "gates/s", "ms_per_step": ms, "steps": steps,
            "shard_gates_this_rank": state["hi"] - state["lo"], "parity_vs_single_rank": "ok",
            "phases_ms_rank0": {k_: v / 2 for k_, v in t_phase.items()},
            "scope": "ONE circuit: stream replayed on every rank (replicated emit), shards planned on the device, each rank builds its gate range, "
                     "one all-gather of counts, wire ids / gate indices shifted to the global numbering; results stay in HBM, sharded by rank",
            "note": "the emit is not sharded (node ids are global prefix counts over the whole stream), so it bounds the speed-up: see phases_ms_rank0"}



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=18315, help="MiMC chains per GPU (546/547 gates each)")
    ap.add_argument("--rounds", type=int, default=91)
    ap.add_argument("--variant", default="late", choices=["inorder", "late"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed end-to-end steps (default = --steps)")
    ap.add_argument("--sample-chains", type=int, default=37, help="chains in the bounded CPU-reference sample (~20 K gates)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-host-emit", action="store_true", help="skip the host-emitter comparison leg")
    ap.add_argument("--stream", default="packed", choices=["packed", "packed6", "aos"],
                    help="event stream format handed to the emitter: packed (kind byte + payload words with implicit operands, ~4 B/event), "
                         "packed6 (every gate / connection operand spelled out, 6 B/event) or 16-byte c2a_event records")
    ap.add_argument("--no-pipelined", action="store_true", help="skip the two-circuits-in-flight e2e leg")
    ap.add_argument("--no-from-source", action="store_true", help="skip the .circom-text-to-circuit leg")
    ap.add_argument("--source-chains", type=int, default=0, help="MiMC chains of the from_source leg (default: --chains, the headline workload)")
    ap.add_argument("--no-phase-timing", action="store_true", help="diagnostic: run the timed loop without the per-kernel CUDA events")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling leg (one circuit sharded across the ranks)")
    ap.add_argument("--legs", default="all", help="comma list of the extra sub-records (N = 1): configs,variants,kahn,sweeps,same_config,deep  (all | none)")
    ap.add_argument("--extra-steps", type=int, default=20, help="timed steps of the small-config sub-records")
    args = ap.parse_args()
    legs = set("configs,variants,kahn,sweeps,same_config,deep".split(",")) if args.legs == "all" else set(x for x in args.legs.split(",") if x and x != "none")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3)
    K = max(args.steps, 1)

    import numpy as np

    workload_name = f"mimc_chains W={args.chains} x {args.rounds} rounds x 6 gates, variant={args.variant}"
    metric = "gates/sec (emit+topo-sort) on >=1M-gate circuit"

    # ------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        # the reference arm never imports the product package: libc2a.so is not mapped into this process (workloads.py is
        # numpy-only and is loaded on its own; the oracle is oracle/libc2a_oracle.so through tests/oracle_lib.py)
        workloads = load_workloads_without_native()
        ncores = os.cpu_count()
        vals = []
        last = None
        for i in range(W + K):
            r = cpu_reference_measure(workloads, None, args.sample_chains, args.variant, args.rounds)
            if i >= W:
                vals.append(r["gates_per_s"])
            last = r
            if i == 0 and (r["emit_s"] + r["backend_s"]) * (W + K) > 240:  # keep the whole run within minutes
                vals = [r["gates_per_s"]]
                break
        v = float(np.mean(vals))
        sample = (f"first {args.sample_chains} chains ({last['sample_gates']} gates) of the workload: faithful O(G*S) emit "
                  f"{last['emit_s']:.2f}s + HashMap/DFS back end {last['backend_s']*1e3:.1f}ms per step; the reference is single-threaded")
        mapped = sorted({ln.split()[-1] for ln in open("/proc/self/maps") if ln.rstrip().endswith(".so") and ROOT in ln})
        assert not any(m.endswith("libc2a.so") for m in mapped), "the reference arm must not map the product library"
        print(json.dumps({
            "impl": "reference", "repo_libs_mapped": [os.path.relpath(m, ROOT) for m in mapped], "metric": metric, "value": v, "unit": "gates/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": W,
            "ms_per_step": 1e3 * last["sample_gates"] / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": {"workload": workload_name, "sample": sample},
            "cpu_baseline": {"value": v, "unit": "gates/s", "cores": 1, "kind": "port", "sample": sample, "host_cores": ncores,
                             "emit_s": last["emit_s"], "backend_s": last["backend_s"]},
            "e2e": {"value": v, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ------------------------------------------------------------------------------------------------
    from c2a_loader import c2a
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this framework has no CPU fallback on the sort path")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        # NCCL writes its version banner / debug lines to STDOUT by default; rank 0's stdout must be the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = c2a.lib
    ctx = c2a.DeviceContext(local_rank)
    h = ctx.handle
    stream = torch.cuda.ExternalStream(lib.c2a_stream(h), device=torch.device("cuda", local_rank))

    # ---- workload (every rank: its own independent component subtree)
    from circom_2_arithc_b200._lib import EmitInfo
    wl = c2a.workloads.mimc_chains(args.chains, rounds=args.rounds, variant=args.variant)
    dev = torch.device("cuda", local_rank)
    from circom_2_arithc_b200._lib import PackedEvents
    packed = args.stream in ("packed", "packed6")
    ev_np = np.ascontiguousarray(wl.events)
    n_ev = int(ev_np.shape[0])
    if packed:   # the walker's output, host side: kinds byte + payload words (c2a_program_packed / c2a_pack_events_ex)
        kinds_np, words_np, pk_flags = c2a.pack_events(ev_np, implicit=(args.stream == "packed"))
        p_kinds = torch.from_numpy(kinds_np).pin_memory()
        p_words = torch.from_numpy(words_np.view(np.int32)).pin_memory()
        d_kinds, d_words = p_kinds.to(dev), p_words.to(dev)
        pk_host = PackedEvents(p_kinds.data_ptr(), p_words.data_ptr(), n_ev, int(words_np.shape[0]), pk_flags, 0)
        pk_dev = PackedEvents(d_kinds.data_ptr(), d_words.data_ptr(), n_ev, int(words_np.shape[0]), pk_flags, 0)
        stream_bytes, stream_bytes_count = n_ev + 4 * int(words_np.shape[0]), n_ev
    else:
        p_events = torch.from_numpy(ev_np.view(np.int32)).pin_memory()
        d_events = p_events.to(dev)
        stream_bytes = stream_bytes_count = 16 * n_ev

    def emit_resident():
        if packed:
            return lib.c2a_emit_packed_resident(h, C.byref(pk_dev), C.byref(info), C.byref(bad))
        return lib.c2a_emit_events_resident(h, vp(d_events.data_ptr()), n_ev, C.byref(info), C.byref(bad))

    def emit_from_host():
        if packed:
            return lib.c2a_emit_packed_device(h, C.byref(pk_host), C.byref(info), C.byref(bad))
        return lib.c2a_emit_events_device(h, vp(p_events.data_ptr()), n_ev, C.byref(info), C.byref(bad))

    in_ids = np.array(sorted(wl.inputs), dtype=np.uint32)
    out_ids = np.array(sorted(wl.outputs), dtype=np.uint32)
    n_const = int(((wl.events[:, 0] & 0xFF) == 1).sum())
    vp = C.c_void_p
    info = EmitInfo()
    bad = C.c_uint64(0)
    wc = C.c_uint32(0)
    err = C.c_uint64(0)

    # sizes (one untimed emit)
    st = emit_resident()
    if st != 0:
        raise RuntimeError(f"emit (resident) -> {st}: {ctx.last_error()}")
    if info.path != 1:
        raise RuntimeError(f"the device emitter declined the workload (flags {info.decline_flags}): nothing to measure")
    G, nb = int(info.n_gates), int(info.node_count) + 1
    d_order = torch.empty(G, dtype=torch.int32, device=dev)
    d_wire = torch.empty(nb, dtype=torch.int32, device=dev)
    d_new = torch.empty((G, 4), dtype=torch.int32, device=dev)
    d_counts = torch.zeros(4, dtype=torch.int64, device=dev)
    d_all = torch.zeros(4 * world, dtype=torch.int64, device=dev)
    phase_acc = {}

    def acc_phases(prefix):
        for k, v in ctx.phases().items():
            phase_acc[prefix + k] = phase_acc.get(prefix + k, 0.0) + v

    h_counts = torch.zeros(4, dtype=torch.int64).pin_memory()

    def reconcile():
        # global wire numbering across ranks: the build above numbered this rank's circuit without gathering the gates; one NCCL
        # all-gather of (n_in, n_mid, n_out, G), then the gather kernel applies the global offsets it derives from the gathered
        # counts on the device - no host read in between, no separate rebase pass, everything is stream-ordered
        h_counts[0], h_counts[1], h_counts[2], h_counts[3] = len(in_ids), wc.value - len(in_ids) - len(out_ids), len(out_ids), G
        with torch.cuda.stream(stream):
            d_counts.copy_(h_counts, non_blocking=True)
            dist.all_gather_into_tensor(d_all, d_counts)
        st = lib.c2a_emitted_gather_device(h, vp(d_order.data_ptr()), vp(d_new.data_ptr()), vp(d_all.data_ptr()), rank, world)
        if st != 0:
            raise RuntimeError(f"c2a_emitted_gather_device -> {st}: {ctx.last_error()}")

    dom_probe = {"name": None, "in_emit": False, "ms": 0.0}   # timed steps: one double read instead of parsing every phase

    from circom_2_arithc_b200._lib import CompileIO
    use_compile = packed   # one call (c2a_compile_packed*): the emit's final status is read with the build's
    # N > 1: the rank's gates are gathered by reconcile() with the global offsets, not by the build
    io_res = CompileIO(in_ids.ctypes.data_as(vp), out_ids.ctypes.data_as(vp), len(in_ids), len(out_ids), vp(d_order.data_ptr()), vp(d_wire.data_ptr()),
                       vp(d_new.data_ptr()) if world == 1 else None, G, nb, 0)

    def device_step(record=False):
        """emit + build with the event stream already resident in HBM; results stay in HBM"""
        if use_compile:
            st = lib.c2a_compile_packed_resident(h, C.byref(pk_dev), C.byref(io_res), C.byref(info), C.byref(wc), C.byref(bad), C.byref(err))
            if st != 0 or info.path != 1:
                raise RuntimeError(f"c2a_compile_packed_resident -> {st} path {info.path}: {ctx.last_error()}")
            if record:
                acc_phases("")     # (the emit's phases carry their "emit:" prefix already)
            elif dom_probe["name"]:
                dom_probe["ms"] += lib.c2a_last_kernel_ms(h, dom_probe["full"])
            if world > 1:
                reconcile()
            return
        st = emit_resident()
        if st != 0 or info.path != 1:
            raise RuntimeError(f"emit (resident) -> {st} path {info.path}: {ctx.last_error()}")
        if record:
            acc_phases("emit:")
        elif dom_probe["in_emit"]:
            dom_probe["ms"] += lib.c2a_last_kernel_ms(h, dom_probe["name"])
        st = lib.c2a_emitted_build_circuit_device(h, in_ids.ctypes.data_as(vp), len(in_ids), out_ids.ctypes.data_as(vp), len(out_ids),
                                                  vp(d_order.data_ptr()), vp(d_wire.data_ptr()), vp(d_new.data_ptr()) if world == 1 else None,
                                                  C.byref(wc), C.byref(err))
        if st != 0:
            raise RuntimeError(f"c2a_emitted_build_circuit_device -> {st}: {ctx.last_error()}")
        if record:
            acc_phases("")
        elif dom_probe["name"] and not dom_probe["in_emit"]:
            dom_probe["ms"] += lib.c2a_last_kernel_ms(h, dom_probe["name"])
        if world > 1:
            reconcile()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mgp = multi_gpu_parity(c2a, torch, dist, ctx, dev, stream, rank, world, args.variant, args.rounds) if world > 1 else None
    if world > 1:   # the parity legs used the handle: put the headline circuit back
        st = emit_resident()
        assert st == 0 and info.path == 1 and int(info.n_gates) == G

    # ---- value: HBM-resident events -> emit -> build, CUDA events on the handle's stream.
    # Per-kernel CUDA events are not free (a pair costs ~4 us of stream time; ~35 phases per step = 0.3 ms at 10 M gates), so:
    #   warm-up steps run with every phase timed and pick the dominant kernel;
    #   the TIMED steps record events around that kernel only (its live duration feeds `roofline`);
    #   K extra steps after the timed region, with every phase timed again, give the per-kernel table.
    for _ in range(W):
        device_step(record=True)
    warm = {k: v for k, v in phase_acc.items() if k.split(":")[-1].startswith("k_")}
    dom_phase = max(warm, key=warm.get)
    phase_acc.clear()
    if args.no_phase_timing:
        lib.c2a_set_timing(h, 0)
    else:
        lib.c2a_set_timing_only(h, dom_phase.split(":")[-1].encode())
    device_step()
    stop = threading.Event()
    clk_lines = []
    th = threading.Thread(target=clocks_sampler, args=(stop, clk_lines, local_rank), daemon=True)
    th.start()
    time.sleep(0.3)
    barrier()
    launches0 = ctx.kernel_launches()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    if not args.no_phase_timing:
        dom_probe.update(name=dom_phase.split(":")[-1].encode(), full=dom_phase.encode(), in_emit=dom_phase.startswith("emit:"), ms=0.0)
    e0.record(stream)
    for _ in range(K):
        device_step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.kernel_launches() - launches0
    dom_live_ms = max(dom_probe["ms"], 0.0) / K
    dom_probe["name"] = None
    phase_acc.clear()
    lib.c2a_set_timing(h, 1)
    lib.c2a_set_timing_only(h, None)
    for _ in range(K):
        device_step(record=True)
    torch.cuda.synchronize()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = world * G / (ms_per_step * 1e-3)
    order_dev = d_order.cpu().numpy().astype(np.uint32)
    gate_base = int(order_dev.min()) if G else 0
    n_identity = bool((order_dev[:4096] - gate_base == np.arange(min(G, 4096), dtype=np.uint32)).all())
    n_mid = int(wc.value) - len(in_ids) - len(out_ids)
    counts = {"G": G, "NB": nb, "n": n_ev, "S": int(info.signal_bound), "C": int(info.n_connections), "Ceff": int(info.n_effective),
              "n_sig": int(info.n_signals), "n_const": n_const, "n_mid": n_mid, "identity": n_identity, "W": (3 * G + 31) // 32,
              "stream_bytes": stream_bytes, "stream_bytes_count": stream_bytes_count, "dense": bool(packed and (pk_flags & 1))}

    # ---- e2e: event stream in PINNED HOST memory -> emit -> build -> result in pinned host memory, all copies inside the timed region.
    #   e2e (headline)   the reference's result shape (BristolCircuit, src/compiler.rs:452-493): the renumbered gates plus the
    #                    wire ids of the input / output / constant signals (c2a_emitted_signal_wires); wire_count
    #   e2e_all_arrays   additionally the sort order and the whole node -> wire map (what the parity tests compare)
    Ke = args.e2e_steps or K
    p_order = torch.empty(G, dtype=torch.int32).pin_memory()
    p_wire = torch.empty(nb, dtype=torch.int32).pin_memory()
    p_new = torch.empty((G, 4), dtype=torch.int32).pin_memory()
    kinds_all = wl.events[:, 0] & 0xFF
    named = np.concatenate([in_ids, out_ids, wl.events[kinds_all == 1, 1]]).astype(np.uint32)
    p_named = torch.from_numpy(named.view(np.int32)).pin_memory()
    p_named_w = torch.empty(len(named), dtype=torch.int32).pin_memory()
    d_named = torch.empty(len(named), dtype=torch.int32, device=dev)
    d_named_w = torch.empty(len(named), dtype=torch.int32, device=dev)

    io_host = CompileIO(in_ids.ctypes.data_as(vp), out_ids.ctypes.data_as(vp), len(in_ids), len(out_ids), None, None, vp(p_new.data_ptr()), G, 0, 0)
    io_host_all = CompileIO(in_ids.ctypes.data_as(vp), out_ids.ctypes.data_as(vp), len(in_ids), len(out_ids), vp(p_order.data_ptr()), vp(p_wire.data_ptr()),
                            vp(p_new.data_ptr()), G, nb, 0)

    def e2e_step(all_arrays):
        if use_compile and world == 1:   # (N > 1: emit, build, named wires and the rebase against the gathered counts are separate calls)
            st = lib.c2a_compile_packed(h, C.byref(pk_host), C.byref(io_host_all if all_arrays else io_host), C.byref(info), C.byref(wc), C.byref(bad), C.byref(err))
            if st != 0 or info.path != 1:
                raise RuntimeError(f"c2a_compile_packed -> {st} path {info.path}: {ctx.last_error()}")
            if not all_arrays:
                st = lib.c2a_emitted_signal_wires(h, vp(p_named.data_ptr()), len(named), vp(p_named_w.data_ptr()))
                if st != 0:
                    raise RuntimeError(f"c2a_emitted_signal_wires -> {st}: {ctx.last_error()}")
            return
        st = emit_from_host()
        if st != 0 or info.path != 1:
            raise RuntimeError(f"emit (host stream) -> {st} path {info.path}: {ctx.last_error()}")
        if world == 1:
            st = lib.c2a_emitted_build_circuit(h, in_ids.ctypes.data_as(vp), len(in_ids), out_ids.ctypes.data_as(vp), len(out_ids),
                                               vp(p_order.data_ptr()) if all_arrays else None, vp(p_wire.data_ptr()) if all_arrays else None,
                                               vp(p_new.data_ptr()), C.byref(wc), C.byref(err))
            if st != 0:
                raise RuntimeError(f"c2a_emitted_build_circuit -> {st}: {ctx.last_error()}")
            if not all_arrays:
                st = lib.c2a_emitted_signal_wires(h, vp(p_named.data_ptr()), len(named), vp(p_named_w.data_ptr()))
                if st != 0:
                    raise RuntimeError(f"c2a_emitted_signal_wires -> {st}: {ctx.last_error()}")
        else:  # results must be rebased to the global numbering before they leave the device
            st = lib.c2a_emitted_build_circuit_device(h, in_ids.ctypes.data_as(vp), len(in_ids), out_ids.ctypes.data_as(vp), len(out_ids),
                                                      vp(d_order.data_ptr()), vp(d_wire.data_ptr()), None, C.byref(wc), C.byref(err))
            if st != 0:
                raise RuntimeError(f"c2a_emitted_build_circuit_device -> {st}: {ctx.last_error()}")
            reconcile()
            if all_arrays:
                st = lib.c2a_rebase_wire_ids_gathered_device(h, vp(d_wire.data_ptr()), nb, vp(d_all.data_ptr()), rank, world)
            else:
                with torch.cuda.stream(stream):
                    d_named.copy_(p_named, non_blocking=True)
                st = lib.c2a_emitted_signal_wires_device(h, vp(d_named.data_ptr()), len(named), vp(d_named_w.data_ptr()))
                if st == 0:
                    st = lib.c2a_rebase_wire_ids_gathered_device(h, vp(d_named_w.data_ptr()), len(named), vp(d_all.data_ptr()), rank, world)
            if st != 0:
                raise RuntimeError(f"named wires / rebase -> {st}: {ctx.last_error()}")
            with torch.cuda.stream(stream):
                if all_arrays:
                    p_order.copy_(d_order, non_blocking=True)
                    p_wire.copy_(d_wire, non_blocking=True)
                else:
                    p_named_w.copy_(d_named_w, non_blocking=True)
                p_new.copy_(d_new, non_blocking=True)
            stream.synchronize()

    def e2e_measure(all_arrays):
        # per-phase CUDA events cost stream time (~0.3 ms per 10 M-gate step): off inside the timed steps, one extra step with them
        # on afterwards fills `last_call_phases_ms`
        lib.c2a_set_timing(h, 0)
        for _ in range(2):
            e2e_step(all_arrays)
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            e2e_step(all_arrays)
        barrier()
        dt = time.perf_counter() - t0
        lib.c2a_set_timing(h, 1)
        e2e_step(all_arrays)
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    lean = True
    dt = e2e_measure(all_arrays=False)
    e2e_phases = {"emit": ctx.phases()} if world > 1 else {"build": ctx.phases()}
    e2e_value = world * G * Ke / dt
    named_w = p_named_w.numpy().astype(np.uint32).copy()
    dt_all = e2e_measure(all_arrays=True)
    # ---- e2e_pipelined (N = 1, extra): two circuits in flight - a second handle on a second host thread - so that the H2D copy
    #      of one step overlaps the D2H copy and the kernels of the other (PCIe is full duplex; every step still copies its own
    #      input and its own result).  Same calls as `e2e`; reported beside it, not instead of it.
    pipe = None
    if world == 1 and not args.no_pipelined:
        ctx2 = c2a.DeviceContext(local_rank)
        lib.c2a_set_timing(ctx2.handle, 0)
        lib.c2a_set_timing(h, 0)
        p_new2 = torch.empty((G, 4), dtype=torch.int32).pin_memory()
        p_named_w2 = torch.empty(len(named), dtype=torch.int32).pin_memory()
        if packed:
            p_kinds2, p_words2 = p_kinds.clone().pin_memory(), p_words.clone().pin_memory()
            pk_host2 = PackedEvents(p_kinds2.data_ptr(), p_words2.data_ptr(), n_ev, int(words_np.shape[0]), pk_flags, 0)
        else:
            p_events2 = p_events.clone().pin_memory()
        errs = []

        stagger = 0.5 * dt / Ke   # start the second circuit half a step later: its H2D then meets the first one's kernels + D2H

        def worker(hh, which, steps):
            inf, b, w, e = EmitInfo(), C.c_uint64(0), C.c_uint32(0), C.c_uint64(0)
            if which:
                time.sleep(stagger)
            try:
                for _ in range(steps):
                    if packed:
                        st = lib.c2a_emit_packed_device(hh, C.byref(pk_host if which == 0 else pk_host2), C.byref(inf), C.byref(b))
                    else:
                        st = lib.c2a_emit_events_device(hh, vp((p_events if which == 0 else p_events2).data_ptr()), n_ev, C.byref(inf), C.byref(b))
                    if st == 0:
                        st = lib.c2a_emitted_build_circuit(hh, in_ids.ctypes.data_as(vp), len(in_ids), out_ids.ctypes.data_as(vp), len(out_ids), None, None,
                                                           vp((p_new if which == 0 else p_new2).data_ptr()), C.byref(w), C.byref(e))
                    if st == 0:
                        st = lib.c2a_emitted_signal_wires(hh, vp(p_named.data_ptr()), len(named), vp((p_named_w if which == 0 else p_named_w2).data_ptr()))
                    if st != 0:
                        errs.append(st)
                        return
            except Exception as ex:  # noqa: BLE001
                errs.append(repr(ex))

        def run_pair(steps):
            ts = [threading.Thread(target=worker, args=(h, 0, steps)), threading.Thread(target=worker, args=(ctx2.handle, 1, steps))]
            for t_ in ts:
                t_.start()
            for t_ in ts:
                t_.join()

        run_pair(2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_pair(Ke)
        torch.cuda.synchronize()
        dtp = time.perf_counter() - t0
        lib.c2a_set_timing(h, 1)
        if errs:
            pipe = {"error": str(errs[:2])}
        else:
            assert torch.equal(p_new, p_new2) and torch.equal(p_named_w, p_named_w2), "the two pipelined handles disagree"
            pipe = {"value": 2 * Ke * G / dtp, "unit": "gates/s", "s_per_step": dtp / (2 * Ke), "in_flight": 2, "steps": 2 * Ke,
                    "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                    "note": "two handles on two host threads, same calls as e2e; each step copies its own input and result"}
        del ctx2
    stop.set()
    th.join(timeout=2)

    # parity spot checks: resident vs host-buffer results; device emitter vs the product's host union-find emitter
    assert np.array_equal(order_dev, p_order.numpy().astype(np.uint32)), "device-resident and host-buffer paths disagree"
    if world == 1:
        ctx._emit_info = {"n_gates": G, "signal_bound": int(info.signal_bound)}
        nos_all = ctx.emitted_fetch(want_gates=False)[1]
        assert np.array_equal(named_w, p_wire.numpy().astype(np.uint32)[nos_all[named]]), "c2a_emitted_signal_wires disagrees with the wire map"
    host_emit = {}
    if not args.no_host_emit:
        t0 = time.perf_counter()
        comp = c2a.Compiler(context=ctx)
        comp.emit_events(wl.events)
        gates_h = comp.gate_array()
        host_emit["emit_s"] = time.perf_counter() - t0
        ins_n, outs_n = comp.signal_nodes(in_ids), comp.signal_nodes(out_ids)
        t1 = time.perf_counter()
        o2, w2, g2, wc2 = ctx.build_circuit(gates_h, comp.node_count + 1, ins_n, outs_n)
        host_emit["build_s"] = time.perf_counter() - t1
        host_emit["gates_per_s"] = G / (time.perf_counter() - t0)
        st = emit_from_host()
        ctx._emit_info = {"n_gates": G, "signal_bound": int(info.signal_bound)}
        gates_d, _ = ctx.emitted_fetch(want_nodes=False)
        assert np.array_equal(gates_d, gates_h), "device emitter and host emitter disagree on the gate vector"
        assert np.array_equal(o2 + np.uint32(gate_base), p_order.numpy().astype(np.uint32)), "host-emitter and device-emitter pipelines disagree on the order"
        del comp
    else:
        gates_h = ins_n = outs_n = None

    # ---- from_source (N = 1, extra): the same workload family as .circom TEXT -> front end (parse + AST walk on one host core,
    #      csrc/c2a_front.cpp) -> packed stream -> device emitter -> build -> renumbered gates + named wires on the host.
    #      What a user of compile() sees; the walk, not the device, sets this number.
    from_source = None
    if world == 1 and not args.no_from_source:
        Ws = max(1, min(args.source_chains or args.chains, args.chains))
        src_text = c2a.workloads.mimc_circom_source(Ws, args.rounds)
        t0 = time.perf_counter()
        dc = c2a.compile(None, source=src_text, emitter="device", context=ctx)
        t1 = time.perf_counter()
        # one untimed pass sizes the pinned result buffers (as in the e2e leg, they exist before the timed region), then the leg is timed
        info_s = ctx.emit_compressed(dc.compressed())   # literal ranges + replay records cross PCIe; the instances are expanded in HBM
        Gs = int(info_s["n_gates"])
        named_s = np.concatenate([dc.input_signals, dc.output_signals, dc._const_signals]).astype(np.uint32)
        ps_new = p_new if Gs == G else torch.empty((max(Gs, 1), 4), dtype=torch.int32).pin_memory()
        ps_named = torch.from_numpy(named_s.view(np.int32)).pin_memory()
        ps_named_w = torch.empty(max(len(named_s), 1), dtype=torch.int32).pin_memory()
        ins_s, outs_s = np.ascontiguousarray(dc.input_signals), np.ascontiguousarray(dc.output_signals)
        wc_s, err_s = C.c_uint32(0), C.c_uint64(0)

        leg_t = {}

        def device_leg():
            ta = time.perf_counter()
            ctx.emit_compressed(dc.compressed())
            tb = time.perf_counter()
            st_ = lib.c2a_emitted_build_circuit(h, ins_s.ctypes.data_as(vp), len(ins_s), outs_s.ctypes.data_as(vp), len(outs_s), None, None,
                                                vp(ps_new.data_ptr()), C.byref(wc_s), C.byref(err_s))
            assert st_ == 0, ctx.last_error()
            tc = time.perf_counter()
            st_ = lib.c2a_emitted_signal_wires(h, vp(ps_named.data_ptr()), len(named_s), vp(ps_named_w.data_ptr()))
            assert st_ == 0, ctx.last_error()
            leg_t.update(emit_compressed_s=tb - ta, build_incl_gates_d2h_s=tc - tb, named_wires_s=time.perf_counter() - tc)

        device_leg()
        td = time.perf_counter()
        device_leg()
        t2 = time.perf_counter()
        dev_s = t2 - td
        assert info_s["path"] == 1 and wc_s.value > 0 and int(ps_named_w[:len(named_s)].max()) < wc_s.value
        from_source = {"value": Gs / ((t1 - t0) + dev_s), "unit": "gates/s", "gates": Gs, "events": int(dc._n_events),
                       "source_bytes": len(src_text), "walk_s": t1 - t0, "device_s": dev_s, "host_threads": 1,
                       "replay_records": int(dc.compressed().n_replays), "device_s_breakdown": dict(leg_t),
                       "note": "mimc_circom_source(W=%d): parse + AST walk (1 host core; a (template, arguments) pair is interpreted twice at most, later "
                               "instances are replay records) = walk_s; c2a_emit_compressed_device (literal ranges + records cross PCIe, the instances are "
                               "expanded in HBM) + c2a_emitted_build_circuit (gates into pinned host memory) + c2a_emitted_signal_wires = device_s" % Ws}
        del dc

    # ---- extra sub-records (N = 1): the other BASELINE configs, the stress variants, the Kahn levels, the sweeps, and the repo arm on
    #      the reference arm's own sample.  Each is measured like the headline (CUDA events on the handle's stream, warm-up first).
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    extra = {}
    if world == 1 and legs:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as orc
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def flush():
            flush_buf.fill_(1)

        Kx = max(3, args.extra_steps)
        t_extra = time.perf_counter()
        if "same_config" in legs:
            rec, _sc = circuit_leg(c2a, torch, ctx, dev, stream, c2a.workloads.mimc_chains(args.sample_chains, rounds=args.rounds, variant=args.variant),
                                   Kx, peak, flush, cpu_backend=True, orc=orc)
            rec["note"] = "the repo arm on exactly the bounded sample the reference arm / cpu_baseline times (same stream, same result arrays)"
            extra["same_config"] = rec
            del _sc
        if "configs" in legs:
            cfgs = {}
            for key, wlx in (("poseidon_shaped", c2a.workloads.poseidon_shaped()), ("sha256_shaped", c2a.workloads.sha256_shaped()),
                             ("keccak_shaped_x1", c2a.workloads.keccak_shaped(1)), ("keccak_shaped_x2", c2a.workloads.keccak_shaped(2))):
                cfgs[key], _sc = circuit_leg(c2a, torch, ctx, dev, stream, wlx, Kx, peak, flush, cpu_backend=True, orc=orc, concurrent=8)
                del _sc
            extra["configs"] = cfgs
        gates_late = nos_late = None
        if legs & {"variants", "kahn", "sweeps"}:
            st = emit_resident()
            assert st == 0 and info.path == 1
            ctx._emit_info = {"n_gates": G, "signal_bound": int(info.signal_bound)}
            gates_late, nos_late = ctx.emitted_fetch()
        if "variants" in legs:
            var = {}
            perm = np.random.RandomState(1).permutation(G)
            shuffled = np.ascontiguousarray(gates_late[perm])
            var["shuffled_10M"], _keep = backend_leg(c2a, torch, ctx, dev, stream, f"{workload_name}, gate vector shuffled (seed 1)", shuffled, nb,
                                                    nos_late[in_ids], nos_late[out_ids], max(2, K // 2), peak, orc=orc, oracle_reps=1)
            del _keep, shuffled, perm
            wl_in = c2a.workloads.mimc_chains(args.chains, rounds=args.rounds, variant="inorder")
            var["inorder_10M"], _sc = circuit_leg(c2a, torch, ctx, dev, stream, wl_in, max(3, K), peak, flush, cpu_backend=True, orc=orc)
            del _sc, wl_in
            extra["variants"] = var
        if "deep" in legs:
            # worst-case shapes of the exact-order sort at 1 M gates: one DFS tree holding every gate (reversed chain through the lh / rh
            # operand: 1 M forward edges), and a fan-in tree emitted root first.  GPU: bounded walks + pointer jumping (k_relax_loop,
            # k_tree_blocks); CPU: the oracle's DFS (the reference recurses 1 M deep here)
            Gd = 1_000_000
            deep = {}
            for nm, slot in (("reversed_chain_lh_1M", 1), ("reversed_chain_rh_1M", 2)):
                gd = np.zeros((Gd, 4), dtype=np.uint32)
                gd[:, 0], gd[:, 3], gd[:, 3 - slot] = 7, 10 + np.arange(Gd), 1
                gd[:, slot] = 10 + np.arange(Gd) + 1
                gd[Gd - 1, slot] = 2
                deep[nm], _k = backend_leg(c2a, torch, ctx, dev, stream, nm, gd, 10 + Gd + 1, [1, 2], [10], 3, peak, orc=orc, oracle_reps=1)
                del _k
            Gt = (1 << 20) - 1
            gi = np.arange(Gt)
            gt_ = np.zeros((Gt, 4), dtype=np.uint32)
            gt_[:, 3] = 10 + gi
            gt_[:, 1] = np.where(2 * gi + 1 < Gt, 10 + 2 * gi + 1, 1)
            gt_[:, 2] = np.where(2 * gi + 2 < Gt, 10 + 2 * gi + 2, 2)
            deep["fan_in_tree_root_first_1M"], _k = backend_leg(c2a, torch, ctx, dev, stream, "complete binary fan-in tree, heap order", gt_, 10 + Gt, [1, 2], [10], 3, peak, orc=orc, oracle_reps=1)
            del _k, gd, gt_
            extra["worst_case_shapes"] = deep
        if "kahn" in legs:
            kh = {"547_levels": kahn_leg(c2a, torch, ctx, dev, f"{workload_name} (gate vector of the headline circuit)", gates_late, nb, 3, peak)}
            w7 = max(1, G // 7)
            wl7 = c2a.workloads.mimc_chains(w7, rounds=1, variant="late")
            k7, w7w, f7 = c2a.pack_events(np.ascontiguousarray(wl7.events))
            i7 = ctx.emit_packed(k7, w7w, f7)
            g7, _ = ctx.emitted_fetch(want_nodes=False)
            kh["7_levels"] = kahn_leg(c2a, torch, ctx, dev, f"mimc_chains W={w7} x 1 round, variant=late", g7, int(i7["node_count"]) + 1, 3, peak)
            del wl7, k7, w7w, g7
            extra["kahn"] = kh
        if "sweeps" in legs:
            kinds_all_ = wl.events[:, 0] & 0xFF
            c_sig, c_val = wl.events[kinds_all_ == 1, 1], wl.events[kinds_all_ == 1, 2]
            t0 = time.perf_counter()
            cm, _cv, dm = ctx.sweep_masks(gates_late, nb, nos_late[c_sig], c_val, nos_late[out_ids])
            dt_sw = time.perf_counter() - t0
            ph = ctx.phases()
            extra["sweeps"] = {"workload": workload_name, "gates": G, "fold_ms": ph.get("k_fold_level"), "dead_ms": ph.get("k_live_level"),
                               "kahn_levels_ms": sum(v for k, v in ph.items() if k.startswith("k_kahn") or k in ("k_producer", "k_deps", "k_level_sort")),
                               "call_s_incl_copies": dt_sw, "const_gates": int(cm.sum()), "dead_gates": int(dm.sum()),
                               "note": "c2a_sweep_masks (K8 constant-fold mask, K9 dead-gate mask) over the Kahn levels; host buffers, copies inside call_s"}
        extra["extras_wall_s"] = time.perf_counter() - t_extra
        del gates_late, nos_late

    strong = None
    if world > 1 and not args.no_strong:
        strong = strong_scaling_leg(c2a, torch, dist, ctx, dev, stream, rank, world, wl, max(3, K // 2))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (CUDA-event time of the phase, measured live above)
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    kern = {k: v / K for k, v in phase_acc.items() if k.split(":")[-1].startswith("k_")}
    dom = dom_phase
    dom_ms = dom_live_ms if dom_live_ms > 0 else kern[dom]
    ab = alg_bytes(dom, counts)
    achieved = ab / (dom_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom.split(":")[-1])
    except Exception:
        pass
    inits = {k: v / K for k, v in phase_acc.items() if k.split(":")[-1] == "init"}
    all_bytes = sum(alg_bytes(k, counts) for k in list(kern) + list(inits))
    roof = {"bound": "hbm", "kernel": dom.split(":")[-1], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0, "alg_bytes_per_launch": ab, "kernel_ms": dom_ms,
            "kernel_share_of_step": dom_ms / ms_per_step,
            "kernel_ms_source": "CUDA events around this kernel inside the timed steps" if dom_live_ms > 0 else "extra steps",
            "per_kernel_note": f"{K} extra steps after the timed region with every phase timed (events around all ~35 phases cost ~0.3 ms/step)",
            "per_kernel_ms": {k: round(v, 5) for k, v in sorted({**kern, **inits}.items())},
            "per_kernel_gbs": {k: round(alg_bytes(k, counts) / (v * 1e-3) / 1e9, 1) for k, v in sorted(kern.items()) if v > 0},
            "whole_step_gbs": all_bytes / (ms_per_step * 1e-3) / 1e9, "whole_step_alg_bytes": all_bytes}

    h2d_all = stream_bytes + 4 * (len(in_ids) + len(out_ids))
    d2h_all = 4 * G + 4 * nb + 16 * G + 4 * 32 * 3
    h2d = h2d_all + (4 * len(named) if lean else 0)
    d2h = (16 * G + 4 * len(named) + 4 * 32 * 3) if lean else d2h_all
    out = {
        "metric": metric, "value": value, "unit": "gates/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name, "gates_per_gpu": int(G), "node_bound": int(nb), "n_inputs": int(len(in_ids)), "n_outputs": int(len(out_ids)),
                   "events_per_gpu": n_ev,
                   "stream_format": ("packed: 1 kind byte per event + u32 payload words (dense signal ids implicit%s), %.2f B/event"
                                     % ("; the operand a gate / connection derives from the signal declared last carries no word" if (pk_flags & 2) else
                                        "; 3 words per gate, 2 per connection", stream_bytes / n_ev)) if packed else "c2a_event records, 16 B/event", "signals_per_gpu": counts["n_sig"], "connections_per_gpu": counts["C"], "effective_merges_per_gpu": counts["Ceff"],
                   "boruvka_rounds": int(info.rounds), "order_is_identity": n_identity,
                   "l2": "inputs larger than L2 (event stream %.0f MB, gate array %.0f MB, node arrays %.0f MB each vs 126 MB L2); no flush" % (stream_bytes / 1e6, 16 * G / 1e6, 4 * nb / 1e6),
                   "value_scope": ("event stream resident in HBM -> c2a_compile_packed_resident = device emitter (scatter, Boruvka MSF, node ids, gate resolve) + build "
                                   "(deps, DFS-order reconstruction, wire numbering, gather) in ONE call, the emit's status read together with the build's; results stay in HBM")
                                  if use_compile else
                                  "event stream resident in HBM -> c2a_emit_packed_resident / c2a_emit_events_resident (device emitter: scatter, Boruvka MSF, node ids, gate resolve) -> "
                                  "c2a_emitted_build_circuit_device (producer map, deps, DFS-order reconstruction, wire numbering, gather); results stay in HBM",
                   "e2e_scope": ("event stream in pinned host memory -> c2a_compile_packed (H2D inside; renumbered gates D2H into pinned host buffers) -> "
                                 "c2a_emitted_signal_wires (named signals H2D, their wires D2H); e2e_all_arrays also copies order and the whole wire map")
                                if use_compile else
                                "event stream in pinned host memory -> c2a_emit_packed_device / c2a_emit_events_device (H2D inside) -> c2a_emitted_build_circuit into pinned host buffers "
                                "(new_gates D2H inside) -> c2a_emitted_signal_wires (named signals H2D, their wires D2H); e2e_all_arrays also copies order and the whole wire map",
                   "numa_node": numa,
                   "oracle_pin": "oracle pinned by the reference's own unit / integration vectors (tests/test_oracle_goldens.py); topological_sort has NO upstream "
                                 "test, so the DFS order is pinned by a line-by-line restatement of src/topological_sort.rs:3-50 only",
                   "parallelism": "1 GPU" if world == 1 else f"{world} GPUs, one independent component subtree (W chains) per rank, NCCL all-gather of wire counts, global offsets applied inside the gate gather"},
        "roofline": roof,
        "e2e": {"value": e2e_value, "unit": "gates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": Ke, "s_per_step": dt / Ke,
                "result": ("renumbered gates + wire ids of the %d input/output/constant signals + wire_count (the reference's BristolCircuit contents)" % len(named))
                          if lean else "order + node->wire map + renumbered gates, rebased to the global numbering",
                "last_call_phases_ms": {k: {kk: round(vv, 3) for kk, vv in v.items()} for k, v in e2e_phases.items()}},
        "e2e_all_arrays": {"value": world * G * Ke / dt_all, "unit": "gates/s", "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                           "s_per_step": dt_all / Ke, "result": "order + whole node->wire map + renumbered gates"},
        "gpu_launches": int(launches),
        "clocks": summarize_clocks(clk_lines),
    }
    for k_, v_ in extra.items():
        out[k_] = v_
    if mgp is not None:
        out["multi_gpu_parity"] = mgp
    if strong is not None:
        out["strong_scaling"] = strong
    if pipe is not None:
        if "value" in pipe:
            pipe["h2d_bytes_per_step"], pipe["d2h_bytes_per_step"] = int(h2d), int(d2h)
        out["e2e_pipelined"] = pipe
    if from_source is not None:
        out["from_source"] = from_source
    if host_emit:
        out["e2e_host_emitter"] = {"value": host_emit["gates_per_s"], "unit": "gates/s", "emit_s": host_emit["emit_s"], "build_s": host_emit["build_s"],
                                   "note": "same circuit through the host union-find emitter (c2a_emit_events + c2a_build_circuit), pageable buffers, 1 step"}
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_measure(c2a.workloads, None, args.sample_chains, args.variant, args.rounds,
                                  backend_gates=(gates_h, ins_n, outs_n) if gates_h is not None else None)
        out["cpu_baseline"] = {
            "value": r["gates_per_s"], "unit": "gates/s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
            "sample": (f"first {args.sample_chains} chains ({r['sample_gates']} gates): faithful O(G*S) emit {r['emit_s']:.2f}s + HashMap/DFS back end "
                       f"{r['backend_s']*1e3:.1f}ms" + (f"; plus the back end alone on the full {r['backend_only_full_gates']} gates" if "backend_only_full_gates" in r else "")),
            "emit_s": r["emit_s"], "backend_s": r["backend_s"]}
        if "backend_only_gates_per_s" in r:
            out["cpu_baseline"]["backend_only_gates_per_s"] = r["backend_only_gates_per_s"]
        if "same_config" in out:   # one like-for-like ratio: both arms on the identical bounded sample
            sc_ = out["same_config"]
            sc_["reference_value"] = r["gates_per_s"]
            sc_["same_config"] = True
            sc_["ratio_value_vs_reference"] = sc_["value"] / r["gates_per_s"]
            sc_["ratio_e2e_vs_reference"] = sc_["e2e"]["value"] / r["gates_per_s"]
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
