"""GPU parity of the single-kernel path (csrc/c2a_fused.cuh, c2a_compile_packed*): emit + build of a circuit of up to ~1 M gates
inside ONE cooperative kernel must give, bit for bit, what the multi-kernel pipeline gives (c2a_emit_packed_* followed by
c2a_emitted_build_circuit*) and what the oracle gives (src/compiler.rs:139-278, 321-464; src/topological_sort.rs:3-50):
same c2a_emit_info, node-id gate vector, signal -> node map, DFS order, wire map, renumbered gates, wire count, and the same
status / event index / `detected at i=` on every stream the reference rejects."""
import ctypes as C

import numpy as np
import pytest

from test_gpu_emit import EV_C, EV_G, EV_S, EV_SC, valid_stream

pytestmark = pytest.mark.gpu

FUSED_MAX = 1 << 22


def run_both(ctx, c2a, ev, ins, outs, expect_fused=True):
    """-> the fused result; asserts it equals the multi-kernel result field by field"""
    k, w, f = c2a.pack_events(np.ascontiguousarray(ev))
    c2a.lib.c2a_set_fused_limits(0, 0)
    cerr = None
    try:
        ci, co, cw, cg, cwc = ctx.compile_packed(k, w, f, ins, outs)
        cgates, cnos = ctx.emitted_fetch()
        assert "k_fused_compile" not in ctx.phases()
    except c2a.CircuitError as e:   # a random stream may hold a dependency cycle (a gate's output merged into its own cone)
        assert e.status == c2a.Status.CYCLIC_DEPENDENCY
        cerr = str(e)
    finally:
        c2a.lib.c2a_set_fused_limits(FUSED_MAX, 0)
    if cerr is not None:
        with pytest.raises(c2a.CircuitError) as ex:
            ctx.compile_packed(k, w, f, ins, outs)
        assert str(ex.value) == cerr and "k_fused_compile" in ctx.phases()
        return None
    fi, fo, fw, fg, fwc = ctx.compile_packed(k, w, f, ins, outs)
    assert ("k_fused_compile" in ctx.phases()) == expect_fused
    fgates, fnos = ctx.emitted_fetch()
    strip = lambda d: {a: b for a, b in d.items() if a != "decline_flags"}
    assert strip(fi) == strip(ci), (fi, ci)
    assert np.array_equal(fgates, cgates), "node-id gate vector differs"
    assert np.array_equal(fnos, cnos), "signal -> node map differs"
    assert np.array_equal(fo, co), f"order differs first at {int(np.argmax(fo != co))}"
    assert np.array_equal(fw, cw), f"wire map differs first at node {int(np.argmax(fw != cw))}"
    assert np.array_equal(fg, cg) and fwc == cwc
    # the 4 B/event form of the same stream (implicit operands) through the fused kernel
    ki, wi, fl = c2a.pack_events(np.ascontiguousarray(ev), implicit=True)
    if fl == 3:
        ii, io_, iw, ig, iwc = ctx.compile_packed(ki, wi, fl, ins, outs)
        assert ("k_fused_compile" in ctx.phases()) == expect_fused and strip(ii) == strip(ci)
        assert np.array_equal(io_, co) and np.array_equal(iw, cw) and np.array_equal(ig, cg) and iwc == cwc
    # named-wire look-ups work on the resident result of the fused call
    sig = np.concatenate([np.asarray(ins, dtype=np.uint32), np.asarray(outs, dtype=np.uint32)])
    if len(sig):
        assert np.array_equal(ctx.emitted_signal_wires(sig), fw[fnos[sig]])
    return fi, fo, fw, fg, fwc, fgates, fnos


def check_oracle(orc, ev, ins, outs, res):
    if res is None:     # cyclic: both paths agreed on the message; the oracle must report the same index
        oc = orc.OracleCompiler()
        oc.emit_events(ev)
        nodes = lambda s: np.array([oc.signal_node(int(x)) for x in s], dtype=np.uint32)
        st, _at, *_ = orc.backend_raw(oc.gate_array(), oc.node_count + 1, nodes(ins), nodes(outs))
        assert st == 1
        return
    fi, fo, fw, fg, fwc, fgates, fnos = res
    oc = orc.OracleCompiler()
    oc.emit_events(ev)
    assert np.array_equal(fgates, oc.gate_array()) and fi["node_count"] == oc.node_count
    nodes = lambda s: np.array([oc.signal_node(int(x)) for x in s], dtype=np.uint32)
    st, _, o_order, o_wire, o_gates, o_wc = orc.backend_raw(fgates, fi["node_count"] + 1, nodes(ins), nodes(outs))
    assert st == 0 and o_wc == fwc
    assert np.array_equal(fo, o_order) and np.array_equal(fw, o_wire) and np.array_equal(fg, o_gates)


@pytest.mark.parametrize("seed", range(16))
def test_random_dense_streams(ctx, c2a, orc, seed):
    rng = np.random.RandomState(7000 + seed)
    n = int(rng.choice([12, 60, 400, 3000, 9000]))
    ev = valid_stream(rng, n, id_order="sequential", shape=["random", "path", "star"][seed % 3], p_redundant=[0.0, 0.15, 0.5][seed % 3])
    kinds = ev[:, 0] & 0xFF
    sigs = ev[kinds <= 1, 1]
    gate_outs = ev[kinds == 2, 3]
    ins = rng.choice(sigs, size=min(5, len(sigs)), replace=False).astype(np.uint32)
    outs = rng.choice(gate_outs, size=min(4, len(gate_outs)), replace=False).astype(np.uint32) if len(gate_outs) else np.zeros(0, np.uint32)
    if seed % 4 == 1 and len(ins) > 1:      # a signal listed twice, and one listed as input AND output (compiler.rs:392-395, 446-449)
        ins = np.concatenate([ins, ins[:1]])
        outs = np.concatenate([outs, ins[1:2]])
    res = run_both(ctx, c2a, ev, ins, outs)
    check_oracle(orc, ev, ins, outs, res)


@pytest.mark.parametrize("name", ["poseidon", "sha256_r6", "keccak1_r2", "keccak2_r1", "mimc_late", "mimc_inorder"])
def test_baseline_shaped_workloads(ctx, c2a, orc, name):
    wl = {"poseidon": lambda: c2a.workloads.poseidon_shaped(), "sha256_r6": lambda: c2a.workloads.sha256_shaped(rounds=6),
          "keccak1_r2": lambda: c2a.workloads.keccak_shaped(1, rounds=2), "keccak2_r1": lambda: c2a.workloads.keccak_shaped(2, rounds=1),
          "mimc_late": lambda: c2a.workloads.mimc_chains(23, rounds=91, variant="late"),
          "mimc_inorder": lambda: c2a.workloads.mimc_chains(9, rounds=30, variant="inorder")}[name]()
    ins, outs = np.array(sorted(wl.inputs), dtype=np.uint32), np.array(sorted(wl.outputs), dtype=np.uint32)
    res = run_both(ctx, c2a, wl.events, ins, outs)
    if wl.n_gates <= 15000:      # the oracle's emit is quadratic
        check_oracle(orc, wl.events, ins, outs, res)
    assert res[0]["n_gates"] == wl.n_gates


def test_full_size_configs_match_the_multi_kernel_path(ctx, c2a):
    """BASELINE configs 3 and 4 at full size (116 K / 387 K gates), fused vs multi-kernel"""
    for wl in (c2a.workloads.sha256_shaped(), c2a.workloads.keccak_shaped(2)):
        ins, outs = np.array(sorted(wl.inputs), dtype=np.uint32), np.array(sorted(wl.outputs), dtype=np.uint32)
        run_both(ctx, c2a, wl.events, ins, outs)


def _decl(n):
    return [(EV_S, i, 0, 0) for i in range(n)]


def test_many_boruvka_rounds_and_tag_wraparound(ctx, c2a, orc):
    rng = np.random.RandomState(3)
    n = 4096
    ev = _decl(n) + [(EV_C, int(k), int(k) + 1, 0) for k in rng.permutation(n - 1)] + [(EV_C, 5, 900, 0), (EV_C, 17, 17, 0)]
    ev = np.asarray(ev, dtype=np.uint32)
    res = run_both(ctx, c2a, ev, [0], [])
    assert res[0]["rounds"] >= 5
    check_oracle(orc, ev, [0], [], res)
    n = 1 << 14   # connections ordered by trailing zeros: one round per bit, more than 7 (the best[] tags wrap)
    k = np.arange(n - 1, dtype=np.int64)
    tz = np.array([((int(x) + 1) & -(int(x) + 1)).bit_length() - 1 for x in k])
    k = k[np.lexsort((k, tz))].astype(np.uint32)
    ev = np.zeros((2 * n - 1, 4), dtype=np.uint32)
    ev[:n, 0], ev[:n, 1] = EV_S, np.arange(n)
    ev[n:, 0], ev[n:, 1], ev[n:, 2] = EV_C, k, k + 1
    res = run_both(ctx, c2a, ev, [3], [])
    assert res[0]["rounds"] > 7


@pytest.mark.parametrize("slot", ["lh", "rh"])
def test_deep_forward_chain(ctx, c2a, orc, slot):
    """gate i reads the output of gate i+1 (emitted later): one DFS tree holding every gate; through the rh operand the
    relaxation needs one queue round per hop - inside the fused kernel that is a grid barrier, not a host round trip"""
    n = 700
    ev = _decl(n + 2)           # s_0 .. s_n chain values, s_{n+1} = k
    nxt = n + 2
    for i in range(n):          # tmp_i = s_{i+1} op k ; connect tmp_i -> s_i
        ev.append((EV_S, nxt, 0, 0))
        ev.append((EV_G | (7 << 8), i + 1, n + 1, nxt) if slot == "lh" else (EV_G | (7 << 8), n + 1, i + 1, nxt))
        ev.append((EV_C, nxt, i, 0))
        nxt += 1
    ev = np.asarray(ev, dtype=np.uint32)
    res = run_both(ctx, c2a, ev, [n, n + 1], [0])
    check_oracle(orc, ev, [n, n + 1], [0], res)
    assert res[1].tolist() == list(range(n - 1, -1, -1))


@pytest.mark.parametrize("slot", ["lh", "rh"])
def test_deep_chain_leaves_the_fused_kernel(ctx, c2a, orc, slot):
    """30 K gates, every one with a forward edge: the fused kernel bounds its walks, recognises the deep DAG and hands the stream to
    the multi-kernel path (pointer jumping); same result either way"""
    n = 30000
    ev = _decl(n + 2)
    nxt = n + 2
    for i in range(n):
        ev.append((EV_S, nxt, 0, 0))
        ev.append((EV_G | (7 << 8), i + 1, n + 1, nxt) if slot == "lh" else (EV_G | (7 << 8), n + 1, i + 1, nxt))
        ev.append((EV_C, nxt, i, 0))
        nxt += 1
    ev = np.asarray(ev, dtype=np.uint32)
    res = run_both(ctx, c2a, ev, [n, n + 1], [0], expect_fused=False)
    assert res[1].tolist() == list(range(n - 1, -1, -1))


def test_cycles_and_self_loops_report_the_reference_index(ctx, c2a, orc):
    for cyc in ("pair", "self", "late"):
        ev = _decl(6)
        if cyc == "pair":       # s1 = s2 + s0 ; s2 = s1 + s0
            body = [(6, 2, 0, 1), (7, 1, 0, 2)]
        elif cyc == "self":     # s1 = s1 + s0
            body = [(6, 0, 3, 4), (7, 1, 0, 1)]
        else:                   # a clean gate first, the cycle behind it
            body = [(6, 0, 3, 4), (7, 2, 0, 1), (8, 1, 0, 2)]
        nxt = 6
        for (t, a, b, o) in body:
            ev.append((EV_S, nxt, 0, 0))
            ev.append((EV_G, a, b, nxt))
            ev.append((EV_C, nxt, o, 0))
            nxt += 1
        ev = np.asarray(ev, dtype=np.uint32)
        k, w, f = c2a.pack_events(ev)
        errs = []
        for lim in (0, FUSED_MAX):
            c2a.lib.c2a_set_fused_limits(lim, 0)
            with pytest.raises(c2a.CircuitError) as ex:
                ctx.compile_packed(k, w, f, [0], [])
            assert ex.value.status == c2a.Status.CYCLIC_DEPENDENCY
            errs.append(str(ex.value))
        c2a.lib.c2a_set_fused_limits(FUSED_MAX, 0)
        oc = orc.OracleCompiler()
        oc.emit_events(ev)
        st, at, *_ = orc.backend_raw(oc.gate_array(), oc.node_count + 1, [oc.signal_node(0)], [])
        assert st == 1 and errs[0] == errs[1] == f"Cyclic dependency: detected at i={at}"


@pytest.mark.parametrize("seed", range(12))
def test_streams_the_reference_rejects_fall_back_to_the_exact_replay(ctx, c2a, orc, seed):
    """dense packed streams with references to signals declared later, merge errors: the fused kernel commits nothing and the
    call returns exactly what c2a_emit_packed_device returns"""
    rng = np.random.RandomState(4000 + seed)
    n_sig_total = int(rng.randint(6, 60))
    ev, declared = [], 0
    while declared < n_sig_total:
        x = rng.rand()
        if x < 0.45 or declared < 3:
            ev.append((EV_SC if rng.rand() < 0.2 else EV_S, declared, 0, 0))
            declared += 1
        elif x < 0.8:
            hi = declared + (3 if rng.rand() < 0.25 else 0)
            a, b = (int(rng.randint(0, min(hi, n_sig_total))) for _ in range(2))
            ev.append((EV_S, declared, 0, 0))
            o = declared if rng.rand() < 0.9 else min(declared + 2, n_sig_total - 1)
            declared += 1
            ev.append((EV_G | (int(rng.randint(0, 20)) << 8), a, b, o))
        else:
            hi = declared + (2 if rng.rand() < 0.2 else 0)
            a, b = (int(rng.randint(0, min(hi, n_sig_total))) for _ in range(2))
            ev.append((EV_C, a, b, 0))
    ev = np.asarray(ev, dtype=np.uint32)
    k, w, f = c2a.pack_events(ev)
    assert f == 1
    oc = orc.OracleCompiler()
    try:
        oc.emit_events(ev)
        err = None
    except orc.OracleError as e:
        err = e
    if err is None:
        try:
            info, order, wire, ng, wc = ctx.compile_packed(k, w, f, [0], [])
        except c2a.CircuitError as e:   # an accepted emit may still hold a dependency cycle
            assert e.status == c2a.Status.CYCLIC_DEPENDENCY
            return
        gates, nos = ctx.emitted_fetch()
        assert np.array_equal(gates, oc.gate_array()) and info["node_count"] == oc.node_count
        st, _, o_order, o_wire, o_gates, o_wc = orc.backend_raw(gates, oc.node_count + 1, [oc.signal_node(0)], [])
        assert st == 0 and np.array_equal(order, o_order) and np.array_equal(wire, o_wire) and np.array_equal(ng, o_gates) and wc == o_wc
    else:
        for imp in (False, True):
            with pytest.raises((c2a.CircuitError, c2a.C2AError)) as ex:
                ctx.compile_packed(*c2a.pack_events(ev, implicit=imp), [0], [])
            assert int(ex.value.status) == err.status
            assert f"event {ex.value.err_event}" == err.message


@pytest.fixture(params=[True, False], ids=["fused", "multi_kernel"])
def either_path(request, c2a):
    """run the test through the single-kernel path and through the multi-kernel one (deferred emit status)"""
    c2a.lib.c2a_set_fused_limits(FUSED_MAX if request.param else 0, 0)
    yield request.param
    c2a.lib.c2a_set_fused_limits(FUSED_MAX, 0)


def test_bad_io_signal_and_capacities(ctx, c2a, either_path):
    wl = c2a.workloads.mimc_chains(3, rounds=5, variant="late")
    k, w, f = c2a.pack_events(np.ascontiguousarray(wl.events))
    with pytest.raises(c2a.C2AError) as ex:
        ctx.compile_packed(k, w, f, [10 ** 6], [])
    assert "never declared" in str(ex.value)
    from circom_2_arithc_b200._lib import CompileIO, EmitInfo, PackedEvents
    ins = np.array(sorted(wl.inputs), dtype=np.uint32)
    outs = np.array(sorted(wl.outputs), dtype=np.uint32)
    ng = np.empty((wl.n_gates - 1, 4), dtype=np.uint32)
    vp = C.c_void_p
    pk = PackedEvents(k.ctypes.data_as(vp), w.ctypes.data_as(vp), k.shape[0], w.shape[0], f, 0)
    io = CompileIO(ins.ctypes.data_as(vp), outs.ctypes.data_as(vp), len(ins), len(outs), None, None, ng.ctypes.data_as(vp), wl.n_gates - 1, 0, 0)
    info, wc, bad, err = EmitInfo(), C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
    st = c2a.lib.c2a_compile_packed(ctx.handle, C.byref(pk), C.byref(io), C.byref(info), C.byref(wc), C.byref(bad), C.byref(err))
    assert st == c2a.Status.INVALID_ARGUMENT and "gates_cap" in ctx.last_error() and info.n_gates == wl.n_gates
    # the handle is idle and usable right away: the same stream with enough room
    _i, order, _w, gates, _wc = ctx.compile_packed(k, w, f, ins, outs)
    assert ("k_fused_compile" in ctx.phases()) == either_path and len(order) == wl.n_gates and gates.shape == (wl.n_gates, 4)


@pytest.mark.parametrize("exact_caps", [True, False])
def test_resident_form_with_device_arrays(ctx, c2a, exact_caps):
    """c2a_compile_packed_resident: stream and result arrays on the device; arrays at least as large as the bounds are written by the
    kernel itself, exactly-sized ones through a device-to-device copy"""
    import torch
    from circom_2_arithc_b200._lib import CompileIO, EmitInfo, PackedEvents
    wl = c2a.workloads.mimc_chains(31, rounds=40, variant="late")
    ins, outs = np.array(sorted(wl.inputs), dtype=np.uint32), np.array(sorted(wl.outputs), dtype=np.uint32)
    k, w, f = c2a.pack_events(np.ascontiguousarray(wl.events))
    _info, r_order, r_wire, r_gates, r_wc = ctx.compile_packed(k, w, f, ins, outs)
    dev = torch.device("cuda", 0)
    dk, dw = torch.from_numpy(k).to(dev), torch.from_numpy(w.view(np.int32)).to(dev)
    G, nb = wl.n_gates, len(r_wire)
    gcap, wcap = (G, nb) if exact_caps else (w.shape[0] // 3 + 5, k.shape[0] + 10)
    d_order = torch.zeros(gcap, dtype=torch.int32, device=dev)
    d_wire = torch.zeros(wcap, dtype=torch.int32, device=dev)
    d_new = torch.zeros((gcap, 4), dtype=torch.int32, device=dev)
    vp = C.c_void_p
    pk = PackedEvents(dk.data_ptr(), dw.data_ptr(), k.shape[0], w.shape[0], f, 0)
    io = CompileIO(ins.ctypes.data_as(vp), outs.ctypes.data_as(vp), len(ins), len(outs), d_order.data_ptr(), d_wire.data_ptr(), d_new.data_ptr(), gcap, wcap, 0)
    info, wc, bad, err = EmitInfo(), C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
    st = c2a.lib.c2a_compile_packed_resident(ctx.handle, C.byref(pk), C.byref(io), C.byref(info), C.byref(wc), C.byref(bad), C.byref(err))
    assert st == 0, ctx.last_error()
    assert "k_fused_compile" in ctx.phases() and info.n_gates == G and wc.value == r_wc
    assert np.array_equal(d_order[:G].cpu().numpy().astype(np.uint32), r_order)
    assert np.array_equal(d_wire[:nb].cpu().numpy().astype(np.uint32), r_wire)
    assert np.array_equal(d_new[:G].cpu().numpy().astype(np.uint32), r_gates)


def test_one_cta_and_full_grid_agree(ctx, c2a):
    """the grid size is a tuning knob (events per CTA): 1 CTA and the whole GPU must produce the same circuit"""
    wl = c2a.workloads.sha256_shaped(rounds=3)
    ins, outs = np.array(sorted(wl.inputs), dtype=np.uint32), np.array(sorted(wl.outputs), dtype=np.uint32)
    k, w, f = c2a.pack_events(np.ascontiguousarray(wl.events))
    res = []
    try:
        for per_cta in (1 << 22, 64):
            c2a.lib.c2a_set_fused_limits(FUSED_MAX, per_cta)
            res.append(ctx.compile_packed(k, w, f, ins, outs))
            assert "k_fused_compile" in ctx.phases()
    finally:
        c2a.lib.c2a_set_fused_limits(FUSED_MAX, 1024)
    for a, b in zip(res[0][1:], res[1][1:]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("fused", [True, False])
def test_host_wire_map_of_exact_size_is_not_overrun(ctx, c2a, fused):
    """c2a_compile_packed with a HOST wire_of_node of exactly node_count + 1 entries on a stream with redundant connections
    (node_count < signals + connections): the call must not write behind the array.  (The multi-kernel path builds while the emit's
    final node count is still in flight - its provisional bound is signals + connections - so the copy-out has to respect wire_cap.)"""
    from circom_2_arithc_b200._lib import CompileIO, EmitInfo, PackedEvents
    wl = c2a.workloads.mimc_chains(7, rounds=9, variant="late")
    ev = np.ascontiguousarray(wl.events).astype(np.uint32)
    conns = ev[(ev[:, 0] & 0xFF) == EV_C]
    again = conns[::2].copy()                       # every other connection once more, operands swapped: already one node each
    again[:, [1, 2]] = again[:, [2, 1]]
    ev = np.ascontiguousarray(np.concatenate([ev, again]))
    ins, outs = np.array(sorted(wl.inputs), dtype=np.uint32), np.array(sorted(wl.outputs), dtype=np.uint32)
    k, w, f = c2a.pack_events(ev)
    c2a.lib.c2a_set_fused_limits(FUSED_MAX if fused else 0, 0)
    try:
        info0, r_order, r_wire, r_gates, r_wc = ctx.compile_packed(k, w, f, ins, outs)
        nc, G = info0["node_count"], info0["n_gates"]
        n_sig, n_conn = info0["n_signals"], info0["n_connections"]
        assert nc < n_sig + n_conn, "the stream must hold redundant connections for this test to bite"
        guard = 0xDEADBEEF
        wire = np.full(nc + 1 + 4096, guard, dtype=np.uint32)
        order = np.full(G + 64, guard, dtype=np.uint32)
        ng = np.full((G + 64, 4), guard, dtype=np.uint32)
        vp = C.c_void_p
        pk = PackedEvents(k.ctypes.data_as(vp), w.ctypes.data_as(vp), k.shape[0], w.shape[0], f, 0)
        io = CompileIO(ins.ctypes.data_as(vp), outs.ctypes.data_as(vp), len(ins), len(outs), order.ctypes.data_as(vp), wire.ctypes.data_as(vp),
                       ng.ctypes.data_as(vp), G, nc + 1, 0)
        info, wc, bad, err = EmitInfo(), C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
        st = c2a.lib.c2a_compile_packed(ctx.handle, C.byref(pk), C.byref(io), C.byref(info), C.byref(wc), C.byref(bad), C.byref(err))
        assert st == 0, ctx.last_error()
        assert ("k_fused_compile" in ctx.phases()) == fused
        assert (wire[nc + 1:] == guard).all(), "wire_of_node was written behind wire_cap"
        assert (order[G:] == guard).all() and (ng[G:] == guard).all()
        assert np.array_equal(wire[:nc + 1], r_wire[:nc + 1]) and np.array_equal(order[:G], r_order) and np.array_equal(ng[:G], r_gates) and wc.value == r_wc
        # one entry short: the reference-facing error, nothing written behind the array either
        wire[:] = guard
        io2 = CompileIO(ins.ctypes.data_as(vp), outs.ctypes.data_as(vp), len(ins), len(outs), None, wire.ctypes.data_as(vp), None, 0, nc, 0)
        st = c2a.lib.c2a_compile_packed(ctx.handle, C.byref(pk), C.byref(io2), C.byref(info), C.byref(wc), C.byref(bad), C.byref(err))
        assert st == c2a.Status.INVALID_ARGUMENT and "wire_cap" in ctx.last_error()
        assert (wire[nc:] == guard).all()
    finally:
        c2a.lib.c2a_set_fused_limits(FUSED_MAX, 0)


def test_irregular_packed_streams_through_both_compile_paths(ctx, c2a, either_path):
    """a gate type out of range, a word count that does not match the kinds, the implicit-operand flag on a declaration: the count pass
    sees them before the scatter has finished - the multi-kernel compile (which sizes its later launches from totals copied out while the
    scatter runs) has to fall back to the ordinary synchronised status and report what c2a_emit_packed_device reports"""
    wl = c2a.workloads.mimc_chains(40, rounds=30, variant="late")
    ins, outs = np.array(sorted(wl.inputs), dtype=np.uint32), np.array(sorted(wl.outputs), dtype=np.uint32)
    for implicit in (False, True):
        k, w, f = c2a.pack_events(np.ascontiguousarray(wl.events), implicit=implicit)
        good = ctx.compile_packed(k, w, f, ins, outs)
        gi = np.flatnonzero((k & 3) == 2)
        cases = []
        bad = k.copy(); bad[gi[len(gi) // 2]] = (bad[gi[len(gi) // 2]] & 0x83) | (25 << 2); cases.append((bad, w))   # gate type 25
        cases.append((k, w[:-1])); cases.append((k, np.concatenate([w, w[:2]])))                                    # word count off
        if implicit:
            si = np.flatnonzero((k & 3) == 0)
            bad = k.copy(); bad[si[7]] |= 0x80; cases.append((bad, w))                                              # flag on a declaration
        for kk, ww in cases:
            errs = []
            for call in (lambda: ctx.emit_packed(kk, ww, f), lambda: ctx.compile_packed(kk, ww, f, ins, outs)):
                with pytest.raises((c2a.C2AError, c2a.CircuitError)) as ex:
                    call()
                errs.append((int(ex.value.status), str(ex.value)))
            assert errs[0] == errs[1], errs
        again = ctx.compile_packed(k, w, f, ins, outs)      # and the handle is fine afterwards
        assert ("k_fused_compile" in ctx.phases()) == either_path
        for a, b in zip(good[1:4], again[1:4]):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("implicit", [False, True])
def test_big_host_stream_arrives_in_chunks(ctx, c2a, implicit):
    """c2a_compile_packed on a host stream of > 4 Mi payload words: the words are copied on a stream of their own, in chunks, and the
    scatter consumes them while they arrive (it polls the running word count) - same circuit as the two-call form, which copies the
    whole stream first"""
    wl = c2a.workloads.mimc_chains(3000, rounds=91, variant="late")     # 1.64 M gates, 6.9 M events: the multi-kernel path
    ins, outs = np.array(sorted(wl.inputs), dtype=np.uint32), np.array(sorted(wl.outputs), dtype=np.uint32)
    k, w, f = c2a.pack_events(np.ascontiguousarray(wl.events), implicit=implicit)
    assert len(w) >= (1 << 22)
    info2 = ctx.emit_packed(k, w, f)
    gates2, nos2 = ctx.emitted_fetch()
    order2, wire2, ng2, wc2 = ctx.emitted_build_circuit(ins, outs)
    for _ in range(2):
        info, order, wire, ng, wc = ctx.compile_packed(k, w, f, ins, outs)
        assert "k_fused_compile" not in ctx.phases()
        gates, nos = ctx.emitted_fetch()
        strip = lambda d: {a: b for a, b in d.items() if a != "decline_flags"}
        assert strip(info) == strip(info2) and wc == wc2
        assert np.array_equal(gates, gates2) and np.array_equal(nos, nos2)
        nb = info["node_count"] + 1
        assert np.array_equal(order, order2) and np.array_equal(wire[:nb], wire2[:nb]) and np.array_equal(ng, ng2)
    # a stream the device emitter declines (an operand far beyond every declared signal: the reference resolves it to node 0,
    # src/compiler.rs:183), big enough for the chunked copy: the scatter's flag is read with the final status, the call falls back to
    # the exact host replay - same outcome as the two-call form
    wbad = w.copy()
    wbad[len(w) // 2] = 0x7FFFFFF0
    outcomes = []
    for call in (lambda: ctx.emit_packed(k, wbad, f), lambda: ctx.compile_packed(k, wbad, f, ins, outs)[0]):
        try:
            i = call()
            g_, n_ = ctx.emitted_fetch(want_nodes=False)
            outcomes.append(("ok", i["path"], i["node_count"], int(g_.sum(dtype=np.uint64))))
        except (c2a.C2AError, c2a.CircuitError) as e:
            outcomes.append(("err", int(e.status), str(e)))
    assert outcomes[0] == outcomes[1], outcomes
