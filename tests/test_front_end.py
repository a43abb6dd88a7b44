"""The circom-subset front end + walker (csrc/c2a_front.cpp) against the reference's own integration tests
(tests/integration.rs:279-475) and the hand-derived goldens of SURVEY.md §4.  CPU only: the emitted event stream is
replayed into the ORACLE (faithful Compiler restatement), built and simulated there."""
import json
import os

import numpy as np
import pytest

import circom_fixtures as fx
import miniwalker as mw


def to_oracle(orc, comp):
    """replay the walker's calls (with names and I/O tags) into the oracle"""
    oc = orc.OracleCompiler()
    oc.emit_events(comp.events)
    kinds = comp.events[:, 0] & 0xFF
    for sid in comp.events[kinds <= 1, 1]:
        oc.set_signal_name(int(sid), comp.signal_name(int(sid)))
    oc.add_inputs({int(s): comp.signal_name(int(s)) for s in comp.input_signals})
    oc.add_outputs({int(s): comp.signal_name(int(s)) for s in comp.output_signals})
    return oc


def run_named(orc, circ, inputs):
    info = circ["info"]
    vals = {info["input_name_to_wire_index"][k]: v for k, v in inputs.items()}
    for ci in info["constants"].values():
        vals[ci["wire_index"]] = int(ci["value"])
    wires = orc.simulate(circ["gates"], circ["wire_count"], vals)
    return {k: wires[w] for k, w in info["output_name_to_wire_index"].items()}


def compile_run(c2a, orc, src, inputs):
    comp = c2a.compile(None, source=src)
    circ = to_oracle(orc, comp).build_circuit()
    return comp, circ, run_named(orc, circ, inputs)


# ---- tests/integration.rs -----------------------------------------------------------------------------------
def test_add_zero(c2a, orc):  # :279-286
    comp, circ, out = compile_run(c2a, orc, fx.ADD_ZERO, {"0.in": 42})
    assert out == {"0.out": 42}
    assert comp.gate_array().tolist() == [[mw.AAdd, 1, 3, 5]] and comp.node_count == 5          # SURVEY §4 golden
    assert circ["info"]["constants"] == {"0.const_signal_0_2": {"value": "0", "wire_index": 1}}


def test_infix_ops(c2a, orc):  # :288-333
    _, circ, out = compile_run(c2a, orc, fx.INFIX_OPS, {f"0.x{i}": i for i in range(6)})
    assert out == {f"0.{n}": e for n, _o, _l, _r, e in mw.INFIX_OUTPUTS}


def test_matrix_element_multiplication(c2a, orc):  # :335-356
    ins = {f"0.{m}[{i}][{j}]": 2 for m in "ab" for i in range(2) for j in range(2)}
    _, _, out = compile_run(c2a, orc, fx.MAT_ELEM_MUL, ins)
    assert out == {f"0.out[{i}][{j}]": 4 for i in range(2) for j in range(2)}


def test_sum(c2a, orc):  # :358-365
    comp, _, out = compile_run(c2a, orc, fx.SUM, {"0.a": 3, "0.b": 5})
    assert out == {"0.out": 8} and comp.gate_array().tolist() == [[mw.AAdd, 1, 2, 5]] and comp.node_count == 5


def test_x_eq_x(c2a, orc):  # :367-374
    assert compile_run(c2a, orc, fx.X_EQ_X, {"0.x": 37})[2] == {"0.out": 1}


def test_out_of_bounds(c2a):  # :376-391 — exact error string
    with pytest.raises(c2a.ProgramError) as e:
        c2a.compile(None, source=fx.INDEX_OUT_OF_BOUNDS)
    assert str(e.value) == "Runtime error: Index out of bounds"


def test_constant_sum(c2a, orc):  # :393-415 — exact maps
    comp, circ, _ = compile_run(c2a, orc, fx.CONSTANT_SUM, {})
    assert circ["info"]["input_name_to_wire_index"] == {}
    assert circ["info"]["constants"] == {"0.const_signal_8_1": {"value": "8", "wire_index": 0}}
    assert comp.gate_array().shape[0] == 0 and comp.node_count == 3


def test_direct_output(c2a, orc):  # :417-441 — exact maps
    _, circ, _ = compile_run(c2a, orc, fx.DIRECT_OUTPUT, {})
    assert circ["info"]["output_name_to_wire_index"] == {"0.out": 0}
    assert circ["info"]["constants"] == {"0.const_signal_42_1": {"value": "42", "wire_index": 0}}


def test_prefix_ops_negative_golden(c2a, orc):  # :455-475 (ignored upstream): input `c` prefix-matches 0.complementC
    comp = c2a.compile(None, source=fx.PREFIX_OPS)
    with pytest.raises(orc.OracleError) as e:
        to_oracle(orc, comp).build_circuit()
    assert "used for both input 0.complement" in e.value.message and "and output 0.complement" in e.value.message
    assert "0.const_signal_0" in [comp.signal_name(int(s)) for s in comp.input_signals]  # the `c` filter also tags constants


def test_under_constrained_compiles_like_the_reference(c2a):  # :443-453 (ignored upstream: "should error" but does not)
    comp = c2a.compile(None, source=fx.UNDER_CONSTRAINED)
    assert comp.node_count == 1 and comp.gate_array().shape[0] == 0


# ---- same calls as the hand-written stand-in walker (tests/miniwalker.py), fixture by fixture --------------------------
@pytest.mark.parametrize("src,fixture", [(fx.ADD_ZERO, mw.fixture_add_zero), (fx.SUM, mw.fixture_sum), (fx.X_EQ_X, mw.fixture_x_eq_x),
                                         (fx.CONSTANT_SUM, mw.fixture_constant_sum), (fx.DIRECT_OUTPUT, mw.fixture_direct_output),
                                         (fx.INFIX_OPS, mw.fixture_infix_ops), (fx.PREFIX_OPS, mw.fixture_prefix_ops),
                                         (fx.ARRAY_ASSIGNMENT, mw.fixture_array_assignment)], ids=lambda x: getattr(x, "__name__", "src"))
def test_same_emission_as_miniwalker(c2a, orc, src, fixture):
    a = orc.OracleCompiler()
    fixture(a)
    comp = c2a.compile(None, source=src)
    b = to_oracle(orc, comp)
    assert a.gate_array().tolist() == b.gate_array().tolist() == comp.gate_array().tolist()
    assert a.nodes() == b.nodes() == comp.nodes()
    assert a.node_count == comp.node_count


def test_array_assignment_golden(c2a, orc):  # callee body first, caller wiring after
    comp, circ, out = compile_run(c2a, orc, fx.ARRAY_ASSIGNMENT, {f"0.a_in[{i}][{j}]": 1 + 2 * i + j for i in range(2) for j in range(2)})
    assert comp.gate_array().tolist() == [[mw.AAdd, 15, 16, 11], [mw.AAdd, 11, 17, 12], [mw.AAdd, 12, 18, 19]] and comp.node_count == 19
    assert comp.nodes()[19]["signals"] == [12, 9, 4]
    assert circ["order"].tolist() == [0, 1, 2] and circ["wire_count"] == 7 and out == {"0.out": 10}
    assert comp.signal_name(5) == "componentA.in[0][0]"   # callee context is named after the template (runtime.rs:75-77)


def test_main_template_argument(c2a, orc):
    comp, circ, out = compile_run(c2a, orc, fx.MAIN_TEMPLATE_ARGUMENT, {"0.in": 7})
    assert out == {"0.out": 107}
    assert "0.const_signal_100_2" in circ["info"]["constants"]


@pytest.mark.parametrize("n,vals,want", [(2, [2, 3], 1), (5, [2, 3, 1, 5, 4], 3), (4, [9, 1, 1, 1], 0)])
def test_argmax_component_arrays_in_loops(c2a, orc, n, vals, want):  # input/circuit.circom, the CLI's default program
    src = fx.ARGMAX.replace("ArgMax(N)", f"ArgMax({n})")
    comp, circ, out = compile_run(c2a, orc, src, {f"0.in[{i}]": v for i, v in enumerate(vals)})
    assert out == {"0.out": want}
    assert comp.gate_array().shape[0] == 11 * n  # 1 comparison + 2 switchers x 5 gates per element
    # loop-body contexts are dropped per iteration (runtime.rs:166-187): the constant `i` is re-created every time,
    # and each Switcher call has its own const_signal_0 for the prefix minus
    names = [comp.signal_name(int(s)) for s in comp.events[(comp.events[:, 0] & 0xFF) == 1, 1]]
    assert names.count("Switcher.const_signal_0") == 2 * n


# ---- language coverage and errors ----------------------------------------------------------------------------------
def test_functions_while_if_and_compound_assignment(c2a, orc):
    src = """
    function fact(k) { var r = 1; while (k > 1) { r *= k; k -= 1; } return r; }
    function pick(a, b) { if (a > b) { return a; } else { return b; } }
    template T(n) {
        signal input x;
        signal output y[n];
        var acc = 0;
        for (var i = 0; i < n; i++) {
            if (i % 2 == 0) { acc += fact(i + 1); } else { acc = pick(acc, 100); }
            y[i] <== x * acc;
        }
    }
    component main = T(4);
    """
    comp, circ, out = compile_run(c2a, orc, src, {"0.x": 3})
    # acc: i=0 -> 1, i=1 -> 100, i=2 -> 106, i=3 -> 106
    assert out == {"0.y[0]": 3, "0.y[1]": 300, "0.y[2]": 318, "0.y[3]": 318}


def test_reversed_arrows_and_signal_initialisers(c2a, orc):
    src = """
    template T() {
        signal input a; signal input b;
        signal t <== a * b;
        signal output o;
        t + 1 ==> o;
    }
    component main {public [a]} = T();
    """
    _, _, out = compile_run(c2a, orc, src, {"0.a": 6, "0.b": 7})
    assert out == {"0.o": 43}


def test_include(c2a, orc, tmp_path):
    (tmp_path / "lib").mkdir()
    (tmp_path / "lib" / "mul.circom").write_text("template Mul() { signal input a; signal input b; signal output c; c <== a * b; }\n")
    (tmp_path / "main.circom").write_text('pragma circom 2.1.0;\ninclude "lib/mul.circom";\ninclude "lib/mul";\n'
                                          "template Top() { signal input p; signal input q; signal output r; component m = Mul();\n"
                                          " m.a <== p; m.b <== q; r <== m.c; }\ncomponent main = Top();\n")
    comp = c2a.compile(str(tmp_path / "main.circom"))
    circ = to_oracle(orc, comp).build_circuit()
    assert run_named(orc, circ, {"0.p": 5, "0.q": 9}) == {"0.r": 45}


@pytest.mark.parametrize("body,text", [
    ("signal input a; signal output b; a === b;", "Statement not implemented"),                        # process.rs:187, README.md:27
    ("signal input a; signal output b; b <== a > 1 ? a : 1;", "Expression not implemented"),            # :310 InlineSwitchOp
    ("signal output b; var v[2] = [1, 2]; b <== v[0];", "Expression not implemented"),                  # :310 ArrayInLine
    ("signal output b; b <== 4294967296;", "Parsing error"),                                            # :302 constants must fit u32
    ("signal output b; var z = 0 - 5; b <== z;", "Operation error: Subtraction underflow"),             # :805-810
    ("signal output b; var z = 1 / 0; b <== z;", "Operation error: Division by zero"),
    ("signal output b; b <== nope(3);", "Undefined function or template"),
    ("signal output b; signal b;", "Runtime error: Item already declared"),
    ("signal output b; assert(1 == 2);", "Runtime error: Assertion failed"),
    ("signal output b; var u; b <== u + 1;", "Empty data item"),
])
def test_error_strings(c2a, body, text):
    with pytest.raises(c2a.ProgramError) as e:
        c2a.compile(None, source="template T() { %s }\ncomponent main = T();" % body)
    assert str(e.value) == text


def test_main_must_be_a_call(c2a):
    with pytest.raises(c2a.ProgramError) as e:
        c2a.compile(None, source="template T() { signal output b; }\ncomponent main = 3;")
    assert str(e.value) == "Main expression not a call"


def test_merge_errors_surface_as_circuit_errors(c2a):  # compiler.rs:239-245 through process.rs `?`
    with pytest.raises(c2a.ProgramError) as e:
        c2a.compile(None, source="template T() { signal input a; signal output b; b <== a + 1; b <== a * 2; }\ncomponent main = T();")
    assert str(e.value) == "Circuit error: Cannot merge output nodes"


def test_report_matches_oracle(c2a, orc):  # compiler.rs:287-319, 503-531
    for src in (fx.ADD_ZERO, fx.ARRAY_ASSIGNMENT, fx.MAT_ELEM_MUL, fx.ARGMAX.replace("ArgMax(N)", "ArgMax(3)")):
        comp = c2a.compile(None, source=src)
        assert comp.generate_circuit_report() == to_oracle(orc, comp).report()
    rep = c2a.compile(None, source=fx.ADD_ZERO).generate_circuit_report()
    assert rep["outputs"] == [{"id": 5, "names": ["0.out"], "value": None}] and rep["value_type"] == "sint"
    # the native report (c2a_circuit_report_json) against the one assembled from nodes() / gate_array() in Python
    from circom_2_arithc_b200.program import _generate_circuit_report_py
    for src in [fx.INFIX_OPS, fx.PREFIX_OPS, c2a.workloads.mimc_circom_source(5, 7)] + [c[0] for c in fx.WALKER_STRESS if c[1] == 0]:
        comp = c2a.compile(None, source=src)
        comp.update_type("sfloat")
        assert comp.generate_circuit_report() == _generate_circuit_report_py(comp)


def test_large_loop_is_linear(c2a):
    """the reference clones the whole context per iteration (runtime.rs:151-159) and scans all nodes per gate; here a
    200 K-gate loop compiles in about a second"""
    import time
    src = "template T(n) { signal input x; signal output y; signal t[n+1]; t[0] <== x; for (var i = 0; i < n; i++) { t[i+1] <== t[i] * t[i] + i; } y <== t[n]; }\ncomponent main = T(100000);"
    t0 = time.time()
    comp = c2a.compile(None, source=src)
    assert comp.gate_array().shape[0] == 200000
    assert time.time() - t0 < 30


# ---- the walker's runtime model on programs the reference fixtures do not reach ------------------------------------
@pytest.mark.parametrize("case", range(len(fx.WALKER_STRESS)))
def test_walker_stress_programs(c2a, case):
    import ctypes as C
    import hashlib
    src, status, err, n_events, n_signals, digest = fx.WALKER_STRESS[case]
    lib = c2a.lib
    p = lib.c2a_program_new()
    try:
        st = lib.c2a_program_compile_source(p, src.encode(), None, None)
        assert st == status
        assert (lib.c2a_program_error(p).decode() if st else "") == err
        n, ns = int(lib.c2a_program_num_events(p)), int(lib.c2a_program_num_signals(p))
        assert (n, ns) == (n_events, n_signals)
        ev = C.string_at(lib.c2a_program_events(p), 16 * n) if n else b""
        names = [lib.c2a_program_signal_name(p, i) for i in range(ns)]
        u32p = lambda q: C.cast(q, C.POINTER(C.c_uint32))
        ins = [u32p(lib.c2a_program_inputs(p))[i] for i in range(lib.c2a_program_num_inputs(p))]
        outs = [u32p(lib.c2a_program_outputs(p))[i] for i in range(lib.c2a_program_num_outputs(p))]
        assert hashlib.sha256(ev + b"|" + b",".join(names) + b"|" + repr((ins, outs)).encode()).hexdigest()[:16] == digest
    finally:
        lib.c2a_program_free(p)


def _walk(c2a, src):
    import ctypes as C
    lib = c2a.lib
    p = lib.c2a_program_new()
    try:
        st = lib.c2a_program_compile_source(p, src.encode(), None, None)
        n, ns = int(lib.c2a_program_num_events(p)), int(lib.c2a_program_num_signals(p))
        ev = C.string_at(lib.c2a_program_events(p), 16 * n) if n else b""
        return st, lib.c2a_program_error(p), ev, [lib.c2a_program_signal_name(p, i) for i in range(ns)]
    finally:
        lib.c2a_program_free(p)


def test_instance_memo_replays_exactly(c2a, monkeypatch):
    """From the second instance on, a (callable, arguments) pair is replayed from the first one's slice of the recorded stream
    with shifted signal ids instead of being interpreted again: calls, ids and names must be those of the full interpretation"""
    sources = [c2a.workloads.mimc_circom_source(30, 91)] + [c[0] for c in fx.WALKER_STRESS]
    for src in sources:
        monkeypatch.delenv("C2A_FRONT_NO_MEMO", raising=False)
        fast = _walk(c2a, src)
        monkeypatch.setenv("C2A_FRONT_NO_MEMO", "1")
        assert _walk(c2a, src) == fast


def test_runaway_recursion_is_a_call_error_not_a_crash(c2a):
    """10 000 nested calls end in ProgramError::CallError (src/program.rs:81-82); the walk runs on its own large stack so that the
    guard is reached before the native stack ends (the reference has no guard: its process aborts)"""
    src = """pragma circom 2.0.0;
function f(n) { var r = 0; if (n > 0) { r = f(n - 1); } return r + 1; }
template A() { signal input a; signal output b; var q = f(300) + f(300) + f(12000); b <== a + q; }
component main = A();
"""
    st, err, _ev, _names = _walk(c2a, src)
    assert st == 113 and err == b"Call error"
    ok = _walk(c2a, src.replace("f(12000)", "f(9000)"))
    assert ok[0] == 0 and len(ok[3]) == 4


def test_instance_memo_at_a_million_gates(c2a, monkeypatch):
    """BASELINE config 5 at the 1 M-gate point as .circom text: the replayed recording (packed stream, constants, names, I/O tags)
    equals the one obtained by interpreting all 1832 x 91 instances, and it is what c2a_pack_events makes of the 16-byte records"""
    import ctypes as C
    from circom_2_arithc_b200._lib import PackedEvents
    lib = c2a.lib
    src = c2a.workloads.mimc_circom_source(1832, 91).encode()

    def walk():
        p = lib.c2a_program_new()
        assert lib.c2a_program_compile_source(p, src, None, None) == 0
        pk = PackedEvents()
        lib.c2a_program_packed(p, C.byref(pk))
        n, nw, nc, ns = int(pk.n_events), int(pk.n_words), int(lib.c2a_program_num_constants(p)), int(lib.c2a_program_num_signals(p))
        arr = lambda ptr, k, ct: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(k,)).copy()
        out = dict(kinds=arr(pk.kinds, n, C.c_uint8), words=arr(pk.words, nw, C.c_uint32), flags=int(pk.flags),
                   const_ids=arr(lib.c2a_program_constant_signals(p), nc, C.c_uint32), const_vals=arr(lib.c2a_program_constant_values(p), nc, C.c_uint32),
                   ins=arr(lib.c2a_program_inputs(p), lib.c2a_program_num_inputs(p), C.c_uint32),
                   outs=arr(lib.c2a_program_outputs(p), lib.c2a_program_num_outputs(p), C.c_uint32),
                   names=[lib.c2a_program_signal_name(p, i) for i in list(range(0, ns, 9973)) + [ns - 1]],
                   events=arr(lib.c2a_program_events(p), 4 * n, C.c_uint32).reshape(n, 4))
        lib.c2a_program_free(p)
        return out

    monkeypatch.delenv("C2A_FRONT_NO_MEMO", raising=False)
    fast = walk()
    monkeypatch.setenv("C2A_FRONT_NO_MEMO", "1")
    slow = walk()
    assert fast["flags"] == slow["flags"] == 1 and fast["names"] == slow["names"]
    for k in ("kinds", "words", "const_ids", "const_vals", "ins", "outs", "events"):
        assert np.array_equal(fast[k], slow[k]), k
    ev = fast["events"]
    assert int((ev[:, 0] & 0xFF == 2).sum()) == 1832 * 547 and len(fast["ins"]) == 1833 and len(fast["outs"]) == 1832
    kinds_b, words, flags = c2a.pack_events(ev)
    assert flags == 1 and np.array_equal(kinds_b, fast["kinds"]) and np.array_equal(words, fast["words"])
    consts = ev[(ev[:, 0] & 0xFF) == 1]
    assert np.array_equal(consts[:, 1], fast["const_ids"]) and np.array_equal(consts[:, 2], fast["const_vals"])


def _compressed_and_packed(c2a, src):
    """(kinds, words) expanded in numpy from the compressed recording the way the device does it - literal ranges first, then the
    replay records generation by generation, unwritten ranges poisoned - and the host-materialised packed stream"""
    import ctypes as C
    from circom_2_arithc_b200._lib import CompressedEvents, PackedEvents, Replay
    lib = c2a.lib
    p = lib.c2a_program_new()
    try:
        assert lib.c2a_program_compile_source(p, src.encode(), None, None) == 0
        cx = CompressedEvents()
        assert lib.c2a_program_compressed(p, C.byref(cx)) == 0
        n, nw, nr = int(cx.n_events), int(cx.n_words), int(cx.n_replays)
        arr = lambda ptr, k, ct: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(k,)).copy() if k else np.zeros(0, ct)
        kinds, words = arr(cx.kinds, n, C.c_uint8), arr(cx.words, nw, C.c_uint32)
        recs = [(r.k_dst, r.k_src, r.k_len, r.w_dst, r.w_src, r.w_len, r.delta, r.gen) for r in (Replay * nr).from_address(cx.replays)] if nr else []
        k_at = w_at = 0
        for kd, ks, kl, wd, ws, wl, _delta, gen in recs:                      # ascending, disjoint, sources before destinations
            assert kd >= k_at and wd >= w_at and ks + kl <= kd and ws + wl <= wd and 1 <= gen <= cx.max_gen
            k_at, w_at = kd + kl, wd + wl
            kinds[kd:kd + kl] = 0xEE                                             # what the walker left unwritten must not matter
            words[wd:wd + wl] = 0xDEADBEEF
        assert k_at <= n and w_at <= nw
        for g in range(1, int(cx.max_gen) + 1):
            snap_k, snap_w = kinds.copy(), words.copy()                          # records of one generation only read older data
            for kd, ks, kl, wd, ws, wl, delta, gen in recs:
                if gen == g:
                    kinds[kd:kd + kl] = snap_k[ks:ks + kl]
                    words[wd:wd + wl] = snap_w[ws:ws + wl] + np.uint32(delta)
        pk = PackedEvents()
        assert lib.c2a_program_packed(p, C.byref(pk)) == 0 and int(pk.n_events) == n and int(pk.n_words) == nw and pk.flags == cx.flags == 1
        return (kinds, words), (arr(pk.kinds, n, C.c_uint8), arr(pk.words, nw, C.c_uint32)), nr, int(cx.max_gen)
    finally:
        lib.c2a_program_free(p)


def test_compressed_recording_expands_to_the_packed_stream(c2a):
    """c2a_program_compressed: replayed instances are records (destination, source, id shift, generation), not copies; expanding
    them generation by generation gives exactly the packed stream c2a_program_packed materialises"""
    nested = [c[0] for c in fx.WALKER_STRESS if c[1] == 0]
    seen_gen = 0
    for src in [c2a.workloads.mimc_circom_source(40, 91), c2a.workloads.mimc_circom_source(3, 2)] + nested:
        (k1, w1), (k2, w2), nr, max_gen = _compressed_and_packed(c2a, src)
        assert np.array_equal(k1, k2) and np.array_equal(w1, w2)
        seen_gen = max(seen_gen, max_gen)
    (k1, _), _, nr, max_gen = _compressed_and_packed(c2a, c2a.workloads.mimc_circom_source(40, 91))
    assert nr == 38 and max_gen == 1            # instances 3..40 of MiMC(91) are one record each
    assert seen_gen >= 2                        # a replayed instance containing replayed instances


def test_signal_names_in_one_call(c2a):
    dev = c2a.compile(None, source=c2a.workloads.mimc_circom_source(5, 7), emitter="device")
    n = int(c2a.lib.c2a_program_num_signals(dev._prog))
    ids = list(range(n)) + [n + 3]                               # an id that does not exist has the empty name
    assert dev.signal_names(ids) == [dev.signal_name(i) for i in range(n)] + [""]
    assert dev.signal_names([]) == [] and dev.signal_names([0]) == ["0.in[0]"]


def test_poseidon_program(c2a, orc, monkeypatch):
    """BASELINE config 2 as a real .circom program (functions for the constants, nested loops, 2-D signal arrays, component arrays
    instantiated under an `if`): 1 413 gates like the synthetic poseidon_shaped() stream, and the circuit computes the permutation"""
    src = c2a.workloads.poseidon_circom_source()
    comp = c2a.compile(None, source=src)
    assert comp.gate_array().shape[0] == 1413 == c2a.workloads.poseidon_shaped().n_gates
    circ = to_oracle(orc, comp).build_circuit()
    for a, b in ((0, 0), (1, 2), (0xFFFFFFFF, 0x9E3779B9), (123456789, 987654321)):
        assert run_named(orc, circ, {"0.in[0]": a, "0.in[1]": b}) == {"0.out": c2a.workloads.poseidon_reference([a, b])}
    monkeypatch.delenv("C2A_FRONT_NO_MEMO", raising=False)
    fast = _walk(c2a, src)                      # Sbox() is interpreted twice and replayed 79 times
    monkeypatch.setenv("C2A_FRONT_NO_MEMO", "1")
    assert _walk(c2a, src) == fast and fast[0] == 0


def test_sha256_compression_program_gives_the_real_digest(c2a, orc, monkeypatch):
    """The SHA-256 compression function (FIPS 180-4) written in the circom subset on the reference's u32 gate arithmetic: front end ->
    calls -> oracle build_circuit -> simulation = hashlib.  (Templates instantiated hundreds of times with a handful of argument
    tuples: the instance memo interprets each pair twice and replays the rest.)"""
    import hashlib
    import struct
    src = c2a.workloads.sha256_circom_source()
    comp = c2a.compile(None, source=src)
    assert 3000 < comp.gate_array().shape[0] < 4000
    circ = to_oracle(orc, comp).build_circuit()
    iv = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]
    for msg in (b"abc", b"", b"circom-2-arithc on B200: gate graph builder"):
        blk = msg + b"\x80" + b"\0" * (55 - len(msg)) + struct.pack(">Q", 8 * len(msg))
        w = list(struct.unpack(">16I", blk))
        ins = {f"0.h[{i}]": iv[i] for i in range(8)}
        ins.update({f"0.w[{i}]": w[i] for i in range(16)})
        out = run_named(orc, circ, ins)
        digest = b"".join(struct.pack(">I", out[f"0.out[{i}]"]) for i in range(8)).hex()
        assert digest == hashlib.sha256(msg).hexdigest()
        assert [out[f"0.out[{i}]"] for i in range(8)] == c2a.workloads.sha256_compress_reference(iv, w)
    monkeypatch.delenv("C2A_FRONT_NO_MEMO", raising=False)
    fast = _walk(c2a, src)
    monkeypatch.setenv("C2A_FRONT_NO_MEMO", "1")
    assert _walk(c2a, src) == fast and fast[0] == 0


def test_examples_are_the_generated_programs(c2a):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name, text in (("sha256_compress", c2a.workloads.sha256_circom_source()), ("poseidon_t3", c2a.workloads.poseidon_circom_source()),
                       ("mimc_chains_48x91", c2a.workloads.mimc_circom_source(48, 91))):
        path = os.path.join(root, "examples", name + ".circom")
        assert open(path).read() == text
        assert c2a.compile(path).gate_array().shape[0] > 1000    # compile_file path (src/program.rs:18-29)


def test_pathological_expression_depth_is_a_parse_error_not_a_crash(c2a):
    """100 000 nested parentheses / a two-million-term chain used to overflow the native stack (parser recursion, resp. the recursion
    of everything that walks or frees a left-deep tree); behind a C ABI that must be `Parsing error`, not a dead host process"""
    head = "pragma circom 2.0.0; template T(){ signal input a; signal output c; c <== "
    tail = "; } component main = T();"
    for body in ("(" * 100000 + "a" + ")" * 100000, "a" + " + a" * 2000000, "a" + "[a" * 50000 + "]" * 50000, "! " * 100000 + "a"):
        with pytest.raises(c2a.ProgramError) as ex:
            c2a.compile(None, source=head + body + tail, emitter="host")
        assert "Parsing error" in str(ex.value)
    # well inside the bound: still compiles
    comp = c2a.compile(None, source=head + "(" * 1500 + "a" + ")" * 1500 + " + a" * 2000 + tail, emitter="host")
    assert comp.gate_array().shape[0] == 2000
