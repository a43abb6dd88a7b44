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
        "emit:k_ev_scatter": c["stream_bytes"] + 16 * ((n + 1023) // 1024) + (ns * 8 + G * (16 + 1) + C * (8 + 4) if c["dense"] else
                                                                             ns * (4 + 8) + G * (16 + 4) + C * (8 + 4 + 4)),
        "emit:k_ev_check_gates": 0 if c["dense"] else G * (16 + 4 + 12 + 1),
        "emit:k_ev_check_conns": 0 if c["dense"] else C * (8 + 4 + 8),
        "emit:k_msf_pick": C * (8 + 16),                          # first round: conn, 2 RED.MIN best (later rounds run on the shrunken list)
        "emit:k_msf_hook": C * (8 + 8 + 4 + 4),                   # first round: conn, 2 best, parent, eff
        "emit:k_scan_u32": 8 * (C // 32 + 1) + 16 * ((n + 1023) // 1024),   # effective-connection bitmap ranks + the two tile-count scans
        "emit:k_ev_nid_edges": C * (4 + 8) + Ceff * (4 + 8),      # conn_sb, conn (+ bitmap words, L2) ; parent chase + atomicMax on the root
        "emit:k_ev_finalize": S * ((0 if c["dense"] else 4) + 8 + 1 + 4 + 4 + 4) + (G + c["n_const"]) * 8,   # sig_t, meta, outmark, parent, id word, nos + screen atomics
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
    except Exception:
        pass
    return None


def cpu_reference_measure(c2a, wl_full, sample_chains, variant, rounds, backend_gates=None):
    """The reference's CPU path, restated (oracle, faithful data structures), single thread like the reference.
    emit: O(G*S) scans (src/compiler.rs:185-195, 219-226, 260-270) on a bounded prefix of the workload;
    back end: HashMap producer map + DFS + first-seen numbering + gather (src/compiler.rs:388-464)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib as orc
    wl = c2a.workloads.mimc_chains(sample_chains, rounds=rounds, variant=variant)
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
    ap.add_argument("--stream", default="packed", choices=["packed", "aos"],
                    help="event stream format handed to the emitter: packed (kinds byte + payload words, c2a_emit_packed_*) or 16-byte c2a_event records")
    ap.add_argument("--no-pipelined", action="store_true", help="skip the two-circuits-in-flight e2e leg")
    ap.add_argument("--no-from-source", action="store_true", help="skip the .circom-text-to-circuit leg")
    ap.add_argument("--source-chains", type=int, default=0, help="MiMC chains of the from_source leg (default: --chains, the headline workload)")
    ap.add_argument("--no-phase-timing", action="store_true", help="diagnostic: run the timed loop without the per-kernel CUDA events")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3)
    K = max(args.steps, 1)

    from c2a_loader import c2a
    import numpy as np

    workload_name = f"mimc_chains W={args.chains} x {args.rounds} rounds x 6 gates, variant={args.variant}"
    metric = "gates/sec (emit+topo-sort) on >=1M-gate circuit"

    # ------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        ncores = os.cpu_count()
        vals = []
        last = None
        for i in range(W + K):
            r = cpu_reference_measure(c2a, None, args.sample_chains, args.variant, args.rounds)
            if i >= W:
                vals.append(r["gates_per_s"])
            last = r
            if i == 0 and (r["emit_s"] + r["backend_s"]) * (W + K) > 240:  # keep the whole run within minutes
                vals = [r["gates_per_s"]]
                break
        v = float(np.mean(vals))
        sample = (f"first {args.sample_chains} chains ({last['sample_gates']} gates) of the workload: faithful O(G*S) emit "
                  f"{last['emit_s']:.2f}s + HashMap/DFS back end {last['backend_s']*1e3:.1f}ms per step; the reference is single-threaded")
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": v, "unit": "gates/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": W,
            "ms_per_step": 1e3 * last["sample_gates"] / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": {"workload": workload_name, "sample": sample},
            "cpu_baseline": {"value": v, "unit": "gates/s", "cores": 1, "kind": "port", "sample": sample, "host_cores": ncores,
                             "emit_s": last["emit_s"], "backend_s": last["backend_s"]},
            "e2e": {"value": v, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ------------------------------------------------------------------------------------------------
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
    packed = args.stream == "packed"
    ev_np = np.ascontiguousarray(wl.events)
    n_ev = int(ev_np.shape[0])
    if packed:   # the walker's output, host side: kinds byte + payload words (c2a_program_packed / c2a_pack_events)
        kinds_np, words_np, pk_flags = c2a.pack_events(ev_np)
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

    def device_step(record=False):
        """emit + build with the event stream already resident in HBM; results stay in HBM"""
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
        dom_probe.update(name=dom_phase.split(":")[-1].encode(), in_emit=dom_phase.startswith("emit:"), ms=0.0)
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

    def e2e_step(all_arrays):
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
        for _ in range(2):
            e2e_step(all_arrays)
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            e2e_step(all_arrays)
        barrier()
        dt = time.perf_counter() - t0
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

        def device_leg():
            ctx.emit_compressed(dc.compressed())
            st_ = lib.c2a_emitted_build_circuit(h, ins_s.ctypes.data_as(vp), len(ins_s), outs_s.ctypes.data_as(vp), len(outs_s), None, None,
                                                vp(ps_new.data_ptr()), C.byref(wc_s), C.byref(err_s))
            assert st_ == 0, ctx.last_error()
            st_ = lib.c2a_emitted_signal_wires(h, vp(ps_named.data_ptr()), len(named_s), vp(ps_named_w.data_ptr()))
            assert st_ == 0, ctx.last_error()

        device_leg()
        td = time.perf_counter()
        device_leg()
        t2 = time.perf_counter()
        dev_s = t2 - td
        assert info_s["path"] == 1 and wc_s.value > 0 and int(ps_named_w[:len(named_s)].max()) < wc_s.value
        from_source = {"value": Gs / ((t1 - t0) + dev_s), "unit": "gates/s", "gates": Gs, "events": int(dc._n_events),
                       "source_bytes": len(src_text), "walk_s": t1 - t0, "device_s": dev_s, "host_threads": 1,
                       "replay_records": int(dc.compressed().n_replays),
                       "note": "mimc_circom_source(W=%d): parse + AST walk (1 host core; a (template, arguments) pair is interpreted twice at most, later "
                               "instances are replay records) = walk_s; c2a_emit_compressed_device (literal ranges + records cross PCIe, the instances are "
                               "expanded in HBM) + c2a_emitted_build_circuit (gates into pinned host memory) + c2a_emitted_signal_wires = device_s" % Ws}
        del dc

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (CUDA-event time of the phase, measured live above)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
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
                   "stream_format": ("packed: 1 kind byte per event + u32 payload words (3 per gate, 2 per connection; dense signal ids implicit), %.2f B/event"
                                     % (stream_bytes / n_ev)) if packed else "c2a_event records, 16 B/event", "signals_per_gpu": counts["n_sig"], "connections_per_gpu": counts["C"], "effective_merges_per_gpu": counts["Ceff"],
                   "boruvka_rounds": int(info.rounds), "order_is_identity": n_identity,
                   "l2": "inputs larger than L2 (event stream %.0f MB, gate array %.0f MB, node arrays %.0f MB each vs 126 MB L2); no flush" % (stream_bytes / 1e6, 16 * G / 1e6, 4 * nb / 1e6),
                   "value_scope": "event stream resident in HBM -> c2a_emit_packed_resident / c2a_emit_events_resident (device emitter: scatter, Boruvka MSF, node ids, gate resolve) -> "
                                  "c2a_emitted_build_circuit_device (producer map, deps, DFS-order reconstruction, wire numbering, gather); results stay in HBM",
                   "e2e_scope": "event stream in pinned host memory -> c2a_emit_packed_device / c2a_emit_events_device (H2D inside) -> c2a_emitted_build_circuit into pinned host buffers "
                                "(new_gates D2H inside) -> c2a_emitted_signal_wires (named signals H2D, their wires D2H); e2e_all_arrays also copies order and the whole wire map",
                   "numa_node": numa,
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
        r = cpu_reference_measure(c2a, None, args.sample_chains, args.variant, args.rounds,
                                  backend_gates=(gates_h, ins_n, outs_n) if gates_h is not None else None)
        out["cpu_baseline"] = {
            "value": r["gates_per_s"], "unit": "gates/s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
            "sample": (f"first {args.sample_chains} chains ({r['sample_gates']} gates): faithful O(G*S) emit {r['emit_s']:.2f}s + HashMap/DFS back end "
                       f"{r['backend_s']*1e3:.1f}ms" + (f"; plus the back end alone on the full {r['backend_only_full_gates']} gates" if "backend_only_full_gates" in r else "")),
            "emit_s": r["emit_s"], "backend_s": r["backend_s"]}
        if "backend_only_gates_per_s" in r:
            out["cpu_baseline"]["backend_only_gates_per_s"] = r["backend_only_gates_per_s"]
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
