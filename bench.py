#!/usr/bin/env python
"""bench.py — gates/sec of the circom-2-arithc flattening hot path (emit + topo-sort / build_circuit).

    python bench.py --gpus N --steps K --warmup W            # this repo's arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (restated oracle)

Workload (BASELINE.json config 5, the >=1M-gate point the metric is quoted on): synthetic MiMC-chain stream,
W independent chains x 91 rounds x 6 gates emitted in the reference walker's order (circom-2-arithc_b200/
workloads.py).  Default W=18315 -> 10.0 M gates; 'late' variant = component inputs wired after the body, so the
DFS post-order is NOT the identity and the full sort path runs.

A step = one pass of the hot path over one circuit:
  value  build_circuit on HBM-resident gates (K1..K7, c2a_build_circuit_device), CUDA events on the handle's stream
  e2e    event stream (host) -> c2a_emit_events (host union-find) -> c2a_get_gates -> c2a_build_circuit with PINNED HOST
         buffers (H2D of the gate array and D2H of order / wire map / new gates inside the timed region)
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

# Algorithmic bytes per gate of each kernel (SURVEY.md §8d accounting: every array element counted
# once per required read and once per required write; nodes/gate measured on the workload).
def alg_bytes(kernel, G, node_bound, n_nonid):
    npg = node_bound / max(G, 1)
    table = {
        "k_producer": 16 + 4,                 # read gate, RED.MAX producer[out]
        "k_deps": 16 + 8 + 8,                 # read gate, 2 producer gathers, write dep pair
        "k_wire_first": 16 + 12,              # read gate (+4 order when sorted), 3 RED.MIN on wire[]
        "k_wire_scan": 16 + 12 + 4 * npg,     # read gate, 3 wire reads, one wire write per numbered node
        "k_gather": 16 + 12 + 16,             # read gate, 3 wire gathers, write new gate
        "k_relax": 8 + 4,                     # read dep pair, r init (+ out-of-order edges, counted separately)
        "k_sizes": 4 + 4,                     # read r, RED.ADD size[r]
        "k_scan_u32": 4 + 4,
        "k_roots": 4 + 8 + 4,                 # read r, read off pair, write order
        "k_tree_dfs": 0,
        "init": 8 * npg,                      # zero producer[], fill wire[]
    }
    extra_order = 4 if n_nonid else 0
    b = table.get(kernel, 0)
    if kernel in ("k_wire_first", "k_wire_scan", "k_gather"):
        b += extra_order
    return b * G


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
    ap.add_argument("--shuffle", type=int, default=0, help="shuffle the gate vector with this seed (stress)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed end-to-end steps (default min(steps,3))")
    ap.add_argument("--sample-chains", type=int, default=37, help="chains in the bounded CPU-reference sample (~20 K gates)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3)
    K = max(args.steps, 1)

    from c2a_loader import c2a
    import numpy as np

    workload_name = f"mimc_chains W={args.chains} x {args.rounds} rounds x 6 gates, variant={args.variant}" + (f", shuffled(seed={args.shuffle})" if args.shuffle else "")
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = c2a.lib
    ctx = c2a.DeviceContext(local_rank)
    h = ctx.handle
    stream = torch.cuda.ExternalStream(lib.c2a_stream(h), device=torch.device("cuda", local_rank))

    # ---- workload (every rank: its own independent component subtree)
    wl = c2a.workloads.mimc_chains(args.chains, rounds=args.rounds, variant=args.variant)
    events = np.ascontiguousarray(wl.events)
    in_ids = np.array(sorted(wl.inputs), dtype=np.uint32)
    out_ids = np.array(sorted(wl.outputs), dtype=np.uint32)

    def emit(ev):
        comp = c2a.Compiler(context=ctx)
        comp.emit_events(ev)
        return comp

    comp = emit(events)
    gates_h = comp.gate_array()
    if args.shuffle:
        gates_h = c2a.workloads.shuffle_gates(gates_h, args.shuffle)
    G = gates_h.shape[0]
    nb = comp.node_count + 1
    ins = comp.signal_nodes(in_ids)
    outs = comp.signal_nodes(out_ids)
    del comp

    dev = torch.device("cuda", local_rank)
    d_gates = torch.from_numpy(gates_h.view(np.int32)).to(dev)
    d_order = torch.empty(G, dtype=torch.int32, device=dev)
    d_wire = torch.empty(nb, dtype=torch.int32, device=dev)
    d_new = torch.empty((G, 4), dtype=torch.int32, device=dev)
    d_counts = torch.zeros(4, dtype=torch.int64, device=dev)
    d_all = torch.zeros(4 * world, dtype=torch.int64, device=dev)
    wc = C.c_uint32(0)
    err = C.c_uint64(0)
    vp = C.c_void_p

    def device_step():
        st = lib.c2a_build_circuit_device(h, vp(d_gates.data_ptr()), G, nb, ins.ctypes.data_as(vp), len(ins), outs.ctypes.data_as(vp), len(outs),
                                          vp(d_order.data_ptr()), vp(d_wire.data_ptr()), vp(d_new.data_ptr()), C.byref(wc), C.byref(err))
        if st != 0:
            raise RuntimeError(f"c2a_build_circuit_device -> {st}: {ctx.last_error()}")
        if world > 1:  # reconcile the global wire numbering: one NCCL all-gather of (n_in, n_mid, n_out)
            n_mid = wc.value - len(ins) - len(outs)
            with torch.cuda.stream(stream):
                d_counts.copy_(torch.tensor([len(ins), n_mid, len(outs), G], dtype=torch.int64), non_blocking=True)
                dist.all_gather_into_tensor(d_all, d_counts)
            stream.synchronize()
            allc = d_all.view(world, 4).cpu().numpy()
            tot_in, tot_mid = int(allc[:, 0].sum()), int(allc[:, 1].sum())
            off_in = int(allc[:rank, 0].sum())
            off_mid = tot_in + int(allc[:rank, 1].sum()) - len(ins)
            off_out = tot_in + tot_mid + int(allc[:rank, 2].sum()) - len(ins) - n_mid
            st = lib.c2a_rebase_wires_device(h, vp(d_new.data_ptr()), vp(d_order.data_ptr()), G, len(ins), n_mid, off_in, off_mid, off_out, int(allc[:rank, 3].sum()))
            if st != 0:
                raise RuntimeError(f"c2a_rebase_wires_device -> {st}: {ctx.last_error()}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: HBM-resident
    for _ in range(W):
        device_step()
    phase_acc = {}
    stop = threading.Event()
    clk_lines = []
    th = threading.Thread(target=clocks_sampler, args=(stop, clk_lines, local_rank), daemon=True)
    th.start()
    time.sleep(0.3)
    barrier()
    launches0 = ctx.kernel_launches()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        device_step()
        for k, v in ctx.phases().items():
            phase_acc[k] = phase_acc.get(k, 0.0) + v
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.kernel_launches() - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = world * G / (ms_per_step * 1e-3)
    n_identity = bool((d_order[:1024].cpu().numpy().astype(np.uint32) == np.arange(min(G, 1024), dtype=np.uint32)).all())

    # ---- e2e: host events -> emit -> gates -> build (pinned host buffers, H2D + D2H inside)
    Ke = args.e2e_steps or min(K, 3)
    p_gates = torch.empty((G, 4), dtype=torch.int32).pin_memory()
    p_order = torch.empty(G, dtype=torch.int32).pin_memory()
    p_wire = torch.empty(nb, dtype=torch.int32).pin_memory()
    p_new = torch.empty((G, 4), dtype=torch.int32).pin_memory()
    emit_s = []

    def e2e_step():
        t0 = time.perf_counter()
        c = lib.c2a_compiler_new()
        bad = C.c_uint64(0)
        st = lib.c2a_emit_events(c, events.ctypes.data_as(vp), events.shape[0], C.byref(bad))
        assert st == 0, st
        lib.c2a_get_gates(c, vp(p_gates.data_ptr()))
        ii = np.empty(len(in_ids), dtype=np.uint32)
        oo = np.empty(len(out_ids), dtype=np.uint32)
        lib.c2a_signal_nodes(c, in_ids.ctypes.data_as(vp), len(in_ids), ii.ctypes.data_as(vp))
        lib.c2a_signal_nodes(c, out_ids.ctypes.data_as(vp), len(out_ids), oo.ctypes.data_as(vp))
        nb_ = lib.c2a_node_count(c) + 1
        emit_s.append(time.perf_counter() - t0)
        st = lib.c2a_build_circuit(h, vp(p_gates.data_ptr()), G, nb_, ii.ctypes.data_as(vp), len(ii), oo.ctypes.data_as(vp), len(oo),
                                   vp(p_order.data_ptr()), vp(p_wire.data_ptr()), vp(p_new.data_ptr()), C.byref(wc), C.byref(err))
        assert st == 0, (st, ctx.last_error())
        lib.c2a_compiler_free(c)

    e2e_step()  # warm
    emit_s.clear()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    e2e_value = world * G * Ke / dt
    backend_ms = ctx.phases().get("total", 0.0)
    stop.set()
    th.join(timeout=2)

    # parity spot check of the resident result against the e2e (host-buffer) result
    assert np.array_equal(d_order.cpu().numpy(), p_order.numpy()), "device-resident and host-buffer paths disagree"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    kern = {k: v / K for k, v in phase_acc.items() if k.startswith("k_")}
    dom = max(kern, key=kern.get)
    dom_ms = kern[dom]
    ab = alg_bytes(dom, G, nb, not n_identity)
    achieved = ab / (dom_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    all_bytes = sum(alg_bytes(k, G, nb, not n_identity) for k in list(kern) + ["init"])
    roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0, "alg_bytes_per_launch": ab, "kernel_ms": dom_ms,
            "kernel_share_of_step": dom_ms / ms_per_step,
            "per_kernel_ms": {k: round(v, 5) for k, v in sorted({**kern, "init": phase_acc.get("init", 0) / K}.items())},
            "per_kernel_gbs": {k: round(alg_bytes(k, G, nb, not n_identity) / (v * 1e-3) / 1e9, 1) for k, v in kern.items() if v > 0},
            "whole_step_gbs": all_bytes / (ms_per_step * 1e-3) / 1e9}

    out = {
        "metric": metric, "value": value, "unit": "gates/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name, "gates_per_gpu": int(G), "node_bound": int(nb), "n_inputs": int(len(ins)), "n_outputs": int(len(outs)),
                   "order_is_identity": n_identity, "events_per_gpu": int(events.shape[0]),
                   "l2": "inputs larger than L2 (gate array %.0f MB + node arrays %.0f MB each vs 126 MB L2); no flush" % (16 * G / 1e6, 4 * nb / 1e6),
                   "value_scope": "c2a_build_circuit_device on HBM-resident gates: producer map, deps, DFS-order reconstruction, wire numbering, gather",
                   "e2e_scope": "host event stream -> c2a_emit_events (host union-find) -> c2a_get_gates -> c2a_build_circuit with pinned host buffers",
                   "parallelism": "1 GPU" if world == 1 else f"{world} GPUs, one independent component subtree (W chains) per rank, NCCL all-gather of wire counts + wire rebase"},
        "roofline": roof,
        "e2e": {"value": e2e_value, "unit": "gates/s", "h2d_bytes_per_step": int(16 * G + 8 * (len(ins) + len(outs)) + 64),
                "d2h_bytes_per_step": int(4 * G + 4 * nb + 16 * G + 16), "steps": Ke, "s_per_step": dt / Ke,
                "emit_s_per_step": float(np.mean(emit_s)), "backend_device_ms_last_step": backend_ms,
                "backend_only_gates_per_s": world * G / max(dt / Ke - float(np.mean(emit_s)), 1e-9)},
        "gpu_launches": int(launches),
        "clocks": summarize_clocks(clk_lines),
    }
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_measure(c2a, None, args.sample_chains, args.variant, args.rounds, backend_gates=(gates_h, ins, outs))
        out["cpu_baseline"] = {
            "value": r["gates_per_s"], "unit": "gates/s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
            "sample": (f"first {args.sample_chains} chains ({r['sample_gates']} gates): faithful O(G*S) emit {r['emit_s']:.2f}s + HashMap/DFS back end "
                       f"{r['backend_s']*1e3:.1f}ms; plus the back end alone on the full {r['backend_only_full_gates']} gates"),
            "emit_s": r["emit_s"], "backend_s": r["backend_s"], "backend_only_gates_per_s": r["backend_only_gates_per_s"]}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
