"""End to end on the GPU: .circom -> front end -> (host or device) emitter -> build_circuit -> evaluate / output files.
Reference: src/main.rs:15-50, tests/integration.rs:257-277 (`simulation_test`)."""
import json
import os

import numpy as np
import pytest

import circom_fixtures as fx
import miniwalker as mw

pytestmark = pytest.mark.gpu


def simulation_test(c2a, ctx, src, inputs):
    """compile -> build_circuit -> run, all through the product (tests/integration.rs:257-277)"""
    comp = c2a.compile(None, source=src, context=ctx)
    circ = comp.build_circuit()
    vals = {circ.info.input_name_to_wire_index[k]: v for k, v in inputs.items()}
    for ci in circ.info.constants.values():
        vals[ci.wire_index] = int(ci.value)
    wires = ctx.evaluate(circ.gate_array, circ.wire_count, vals)
    return {k: wires[w] for k, w in circ.info.output_name_to_wire_index.items()}, circ, comp


def test_reference_integration_suite(c2a, ctx):
    assert simulation_test(c2a, ctx, fx.ADD_ZERO, {"0.in": 42})[0] == {"0.out": 42}
    assert simulation_test(c2a, ctx, fx.INFIX_OPS, {f"0.x{i}": i for i in range(6)})[0] == {f"0.{n}": e for n, _o, _l, _r, e in mw.INFIX_OUTPUTS}
    ins = {f"0.{m}[{i}][{j}]": 2 for m in "ab" for i in range(2) for j in range(2)}
    assert simulation_test(c2a, ctx, fx.MAT_ELEM_MUL, ins)[0] == {f"0.out[{i}][{j}]": 4 for i in range(2) for j in range(2)}
    assert simulation_test(c2a, ctx, fx.SUM, {"0.a": 3, "0.b": 5})[0] == {"0.out": 8}
    assert simulation_test(c2a, ctx, fx.X_EQ_X, {"0.x": 37})[0] == {"0.out": 1}
    _, circ, _ = simulation_test(c2a, ctx, fx.CONSTANT_SUM, {})
    assert {k: (v.value, v.wire_index) for k, v in circ.info.constants.items()} == {"0.const_signal_8_1": ("8", 0)}
    _, circ, _ = simulation_test(c2a, ctx, fx.DIRECT_OUTPUT, {})
    assert circ.info.output_name_to_wire_index == {"0.out": 0}
    with pytest.raises(c2a.CircuitError) as e:
        simulation_test(c2a, ctx, fx.PREFIX_OPS, {})
    assert "used for both input 0.complement" in str(e.value)
    out, _, _ = simulation_test(c2a, ctx, fx.ARGMAX.replace("ArgMax(N)", "ArgMax(5)"), {f"0.in[{i}]": v for i, v in enumerate([2, 3, 1, 5, 4])})
    assert out == {"0.out": 3}


def test_walker_stream_through_the_device_emitter(c2a, ctx, orc):
    """a loop-heavy program: the walker's event stream replayed on the GPU gives the host emitter's circuit"""
    src = ("template Sq() { signal input a; signal output b; b <== a * a + 1; }\n"
           "template T(n) { signal input x[n]; signal output y[n]; component s[n];\n"
           " for (var i = 0; i < n; i++) { s[i] = Sq(); s[i].a <== x[i] + 1; y[i] <== s[i].b * x[i]; } }\ncomponent main = T(3000);")
    comp = c2a.compile(None, source=src, context=ctx)
    info = ctx.emit_events(comp.events)
    assert info["path"] == 1 and info["n_gates"] == 12000
    gates, nos = ctx.emitted_fetch()
    assert np.array_equal(gates, comp.gate_array()) and info["node_count"] == comp.node_count
    order, wire, ng, wc = ctx.emitted_build_circuit(comp.input_signals, comp.output_signals)
    circ = comp.build_circuit()
    assert np.array_equal(ng, circ.gate_array) and wc == circ.wire_count and np.array_equal(order, circ.order)
    assert not np.array_equal(order, np.arange(12000))  # component bodies precede their input wiring: real reordering
    vals = {int(wire[nos[s]]): 2 + (int(s) % 7) for s in comp.input_signals}
    for k, ci in circ.info.constants.items():
        vals[ci.wire_index] = int(ci.value)
    got = ctx.evaluate(ng, wc, vals)
    for i in range(3000):
        x = 2 + (i % 7)
        assert got[circ.info.output_name_to_wire_index[f"0.y[{i}]"]] == ((x + 1) * (x + 1) + 1) * x


def test_cli_writes_the_three_files(c2a, orc, tmp_path):  # src/main.rs:34-47
    src = tmp_path / "circuit.circom"
    src.write_text(fx.ARGMAX.replace("ArgMax(N)", "ArgMax(2)"))
    out = tmp_path / "out"
    assert c2a.cli_main(["-i", str(src), "-o", str(out), "-v", "sfloat"]) == 0
    info = json.loads((out / "circuit_info.json").read_text())
    report = json.loads((out / "report.json").read_text())
    lines = (out / "circuit.txt").read_text().split("\n")
    comp = c2a.compile(str(src))
    oc = orc.OracleCompiler()
    oc.emit_events(comp.events)
    kinds = comp.events[:, 0] & 0xFF
    for sid in comp.events[kinds <= 1, 1]:
        oc.set_signal_name(int(sid), comp.signal_name(int(sid)))
    oc.add_inputs({int(s): comp.signal_name(int(s)) for s in comp.input_signals})
    oc.add_outputs({int(s): comp.signal_name(int(s)) for s in comp.output_signals})
    want = oc.build_circuit()
    assert info == want["info"]
    assert report == oc.report("sfloat") and report["value_type"] == "sfloat"
    G = want["gates"].shape[0]
    assert lines[0] == f"{G} {want['wire_count']}" and lines[3] == ""
    names = [t.name for t in c2a.AGateType]
    assert lines[4:4 + G] == [f"2 1 {a} {b} {o} {names[op]}" for op, a, b, o in want["gates"].tolist()]




def test_mimc_circom_through_packed_stream_named_wires_and_evaluator(c2a, ctx):
    """BASELINE config 5 as a real .circom program (MiMC-7 rounds x^7 with c_i = i, u32 arithmetic): front end -> packed stream ->
    device emitter -> build (gates only) -> named-wire lookup -> GPU evaluator, checked against the function computed in Python"""
    comp = c2a.compile(None, source=c2a.workloads.mimc_circom_source(48, 91), context=ctx)
    kinds_b, words, flags = c2a.pack_events(comp.events)
    assert flags == 1                                     # the walker numbers its signals densely (src/runtime.rs:120-125)
    info = ctx.emit_packed(kinds_b, words, flags)
    assert info["path"] == 1 and info["n_gates"] == comp.gate_array().shape[0] and info["node_count"] == comp.node_count
    order, wire, ng, wc = ctx.emitted_build_circuit(comp.input_signals, comp.output_signals, want_order=False, want_wires=False)
    assert order is None and wire is None
    ev = comp.events
    const_sigs = ev[(ev[:, 0] & 0xFF) == 1]
    in_w = ctx.emitted_signal_wires(comp.input_signals)
    out_w = ctx.emitted_signal_wires(comp.output_signals)
    const_w = ctx.emitted_signal_wires(const_sigs[:, 1])
    circ = comp.build_circuit()                           # host emitter + c2a_build_circuit on the same handle (drops the resident circuit)
    assert np.array_equal(ng, circ.gate_array) and wc == circ.wire_count
    names_in = [comp.signal_name(int(s)) for s in comp.input_signals]
    assert {n: int(w) for n, w in zip(names_in, in_w)} == circ.info.input_name_to_wire_index
    M = 0xFFFFFFFF
    key, xs = 0x9E3779B9, [(7919 * (w + 1)) & M for w in range(48)]
    vals = {}
    for n, w in zip(names_in, in_w):
        vals[int(w)] = key if n == "0.key" else xs[int(n[len("0.in["):-1])]
    for (_, _sid, v, _), w in zip(const_sigs.tolist(), const_w.tolist()):
        vals[int(w)] = int(v)
    got = ctx.evaluate(ng, wc, vals)

    def mimc(x):
        for i in range(91):
            t = (x + key + i) & M
            t2 = t * t & M
            t4 = t2 * t2 & M
            x = (t4 * t2 & M) * t & M
        return (x + key) & M

    names_out = [comp.signal_name(int(s)) for s in comp.output_signals]
    for n, w in zip(names_out, out_w):
        assert got[int(w)] == mimc(xs[int(n[len("0.out["):-1])]), n


def _same_circuit(a, b):
    assert a.wire_count == b.wire_count
    assert np.array_equal(a.gate_array, b.gate_array) and np.array_equal(a.order, b.order)
    assert a.info.input_name_to_wire_index == b.info.input_name_to_wire_index
    assert a.info.output_name_to_wire_index == b.info.output_name_to_wire_index
    assert {k: (v.value, v.wire_index) for k, v in a.info.constants.items()} == {k: (v.value, v.wire_index) for k, v in b.info.constants.items()}
    assert list(a.info.constants) == list(b.info.constants)  # same (sorted) key order as the JSON of the host path


def test_compile_with_the_device_emitter_gives_the_same_bristol_circuit(c2a, ctx):
    """compile(emitter="device"): the walk only records its calls, build_circuit() replays them on the GPU (packed stream) and
    looks the named wires up - same BristolCircuit and same CircuitError as the host-emitter Compiler (src/compiler.rs:321-494)"""
    sources = [fx.ADD_ZERO, fx.INFIX_OPS, fx.MAT_ELEM_MUL, fx.SUM, fx.X_EQ_X, fx.CONSTANT_SUM, fx.DIRECT_OUTPUT,
               fx.ARGMAX.replace("ArgMax(N)", "ArgMax(5)"), c2a.workloads.mimc_circom_source(5, 7), c2a.workloads.poseidon_circom_source()]
    for src in sources:
        host = c2a.compile(None, source=src, context=ctx).build_circuit()
        dev = c2a.compile(None, source=src, context=ctx, emitter="device")
        _same_circuit(dev.build_circuit(), host)
        assert dev.emit_info["path"] == 1
    with pytest.raises(c2a.CircuitError) as e_host:
        c2a.compile(None, source=fx.PREFIX_OPS, context=ctx).build_circuit()
    with pytest.raises(c2a.CircuitError) as e_dev:
        c2a.compile(None, source=fx.PREFIX_OPS, context=ctx, emitter="device").build_circuit()
    assert str(e_dev.value) == str(e_host.value) and "used for both input 0.complement" in str(e_dev.value)
    with pytest.raises(c2a.ProgramError):
        c2a.compile(None, source="template T() { signal input a; a === 1; } component main = T();", emitter="device")


def test_compressed_recording_is_expanded_on_the_device(c2a, ctx):
    """c2a_emit_compressed_device: only the literal ranges and the replay records of the walker's recording cross PCIe, the replayed
    instances are expanded in HBM generation by generation - same emit_info, gates and signal -> node map as the host-materialised
    packed stream through c2a_emit_packed_device"""
    import ctypes as C
    from circom_2_arithc_b200._lib import CompressedEvents, Replay
    import circom_fixtures as fx
    lib = c2a.lib
    sources = [c2a.workloads.mimc_circom_source(48, 91), c2a.workloads.mimc_circom_source(400, 91), c2a.workloads.mimc_circom_source(1, 1),
               c2a.workloads.poseidon_circom_source(), fx.ADD_ZERO, fx.INFIX_OPS] + [c[0] for c in fx.WALKER_STRESS if c[1] == 0]
    deep = 0
    for src in sources:
        dev = c2a.compile(None, source=src, context=ctx, emitter="device")
        cx = dev.compressed()
        deep = max(deep, int(cx.max_gen))
        info_c = ctx.emit_compressed(cx)
        gates_c, nos_c = ctx.emitted_fetch()
        info_p = ctx.emit_packed(dev._kinds, dev._words, dev._flags)            # materialises the records on the host
        gates_p, nos_p = ctx.emitted_fetch()
        assert info_c == info_p and info_c["path"] == 1
        assert np.array_equal(gates_c, gates_p) and np.array_equal(nos_c, nos_p)
    assert deep >= 2
    # inconsistent records are refused before anything is launched
    dev = c2a.compile(None, source=c2a.workloads.mimc_circom_source(8, 5), context=ctx, emitter="device")
    cx = dev.compressed()
    recs = (Replay * int(cx.n_replays)).from_address(cx.replays)
    bad = (Replay * int(cx.n_replays))(*recs)
    bad[0].k_src = bad[0].k_dst                                                 # a source that overlaps its destination
    cx2 = CompressedEvents(cx.kinds, cx.words, cx.n_events, cx.n_words, C.cast(bad, C.c_void_p), cx.n_replays, cx.max_gen, cx.flags)
    with pytest.raises((c2a.C2AError, c2a.CircuitError)):
        ctx.emit_compressed(cx2)
    # ... and so is a generation that is not larger than the generation of a record its source reads (it would expand to a
    # plausible but wrong stream: bytes unwritten in this launch, or stale ones of the previous call)
    for src in sources:
        dev = c2a.compile(None, source=src, context=ctx, emitter="device")
        cx = dev.compressed()
        if cx.max_gen >= 2:
            break
    assert cx.max_gen >= 2
    recs = (Replay * int(cx.n_replays)).from_address(cx.replays)
    deep_i = max(range(int(cx.n_replays)), key=lambda i: recs[i].gen)
    bad = (Replay * int(cx.n_replays))(*recs)
    bad[deep_i].gen = 1
    cx3 = CompressedEvents(cx.kinds, cx.words, cx.n_events, cx.n_words, C.cast(bad, C.c_void_p), cx.n_replays, cx.max_gen, cx.flags)
    with pytest.raises((c2a.C2AError, c2a.CircuitError)) as ex:
        ctx.emit_compressed(cx3)
    assert "generation" in str(ex.value)
    ctx.emit_compressed(cx)   # the untouched records still expand
