"""Emission ORDER pinned independently of this repo's walker: for three programs with nested components, component arrays in a
loop and forward edges, the add_signal / add_gate / add_connection calls below were derived BY HAND from the reference's rules
(src/process.rs:54-110 declarations, :192-277 substitutions, :315-419 calls - the callee body is emitted at the call site, before
the caller wires the component's inputs -, :426-478 infix operations: lhs constant, rhs constant, output signal, gate; :558-579
constants: one signal per value per context, loop-body contexts are dropped per iteration, src/runtime.rs:151-187; signal ids from
one counter, src/runtime.rs:120-125).  The third program is the LessThan / Num2Bits shape of
tests/circuits/machine-learning/circomlib/comparators.circom:85-87 reduced to the accepted subset.

Asserted for: the front end (csrc/c2a_front.cpp) - its recorded calls must equal the hand-derived list exactly; the oracle, the host
emitter and (GPU) the device emitter - replaying the list must give the same node-id gate vector; for the first program the node ids,
the DFS order and the built circuit were derived by hand as well (src/compiler.rs:139-278, 321-494)."""
import numpy as np
import pytest

S, SC, G, C = 0, 1, 2, 3
AAdd, AMul, ASub, AShiftR, ABitAnd = 0, 7, 9, 15, 19


def sig(i): return (S, i, 0, 0)
def const(i, v): return (SC, i, v, 0)
def gate(op, a, b, o): return (G | (op << 8), a, b, o)
def conn(a, b): return (C, a, b, 0)


P1_SRC = """pragma circom 2.0.0;
template Inner() { signal input x; signal output y; y <== x * x; }
template Main() {
    signal input a; signal output b;
    component c = Inner();
    c.x <== a + 1;
    b <== c.y;
}
component main = Main();
"""
# a=0 b=1 | Inner: x=2 y=3 tmp=4 | const_1=5 tmp=6
P1_EVENTS = [sig(0), sig(1), sig(2), sig(3), sig(4), gate(AMul, 2, 2, 4), conn(4, 3),
             const(5, 1), sig(6), gate(AAdd, 0, 5, 6), conn(6, 2), conn(3, 1)]
# nodes: signals 0..4 -> 1..5; conn(4,3) -> 6; const 5 -> 7; tmp 6 -> 8; conn(6,2) -> 9 = {6,2}; conn(3,1) -> 10 = {4,3,1}
P1_GATES = [[AMul, 9, 9, 10], [AAdd, 1, 7, 9]]      # gate 0 reads node 9, which gate 1 (emitted later) produces: a forward edge
P1_NODE_COUNT = 10
P1_ORDER = [1, 0]                                    # DFS from root 0 visits its dependency gate 1 first
P1_NEW_GATES = [[AAdd, 0, 1, 2], [AMul, 2, 2, 3]]    # wires: a -> 0, const_1 -> 1 (first seen), node 9 -> 2, b -> 3 (outputs last)
P1_WIRE_COUNT = 4

P2_SRC = """pragma circom 2.0.0;
template Sq() { signal input x; signal output y; y <== x * x; }
template Twice() { signal input p; signal output q; component s = Sq(); s.x <== p + 2; q <== s.y + p; }
template Main() {
    signal input in[2]; signal output out[2];
    component t[2];
    for (var i = 0; i < 2; i++) { t[i] = Twice(); t[i].p <== in[i]; out[i] <== t[i].q; }
}
component main = Main();
"""


def _p2_instance(base, k):   # Twice: p q | Sq: x y tmp | const_2 tmp | tmp
    p, q, x, y, t0, c2, t1, t2 = (base + j for j in range(8))
    return [sig(p), sig(q), sig(x), sig(y), sig(t0), gate(AMul, x, x, t0), conn(t0, y),
            const(c2, 2), sig(t1), gate(AAdd, p, c2, t1), conn(t1, x),
            sig(t2), gate(AAdd, y, p, t2), conn(t2, q),
            conn(k, p), conn(q, 2 + k)]                       # caller wiring AFTER the body: in[k] -> t[k].p, t[k].q -> out[k]


P2_EVENTS = [sig(0), sig(1), sig(2), sig(3)] + _p2_instance(4, 0) + _p2_instance(12, 1)

P3_SRC = """pragma circom 2.0.0;
template Bits(n) { signal input in; signal output out[n]; for (var i = 0; i < n; i++) { out[i] <== (in >> i) & 1; } }
template LessThan(n) {
    signal input in[2]; signal output out;
    component b = Bits(n + 1);
    b.in <== in[0] + (1 << n) - in[1];
    out <== 1 - b.out[n];
}
component main = LessThan(2);
"""
P3_EVENTS = [sig(0), sig(1), sig(2), sig(3), sig(4), sig(5), sig(6),
             # i = 0: fresh loop-body context: const_0 (the value of i), then const_1
             const(7, 0), sig(8), gate(AShiftR, 3, 7, 8), const(9, 1), sig(10), gate(ABitAnd, 8, 9, 10), conn(10, 4),
             # i = 1: the value of i IS 1: const_signal_1 is created once and reused by `& 1` in the same context
             const(11, 1), sig(12), gate(AShiftR, 3, 11, 12), sig(13), gate(ABitAnd, 12, 11, 13), conn(13, 5),
             # i = 2
             const(14, 2), sig(15), gate(AShiftR, 3, 14, 15), const(16, 1), sig(17), gate(ABitAnd, 15, 16, 17), conn(17, 6),
             # caller: (in[0] + (1 << 2)) - in[1] -> b.in   (1 << n is folded on the host: src/process.rs:440-452)
             const(18, 4), sig(19), gate(AAdd, 0, 18, 19), sig(20), gate(ASub, 19, 1, 20), conn(20, 3),
             # out <== 1 - b.out[2]
             const(21, 1), sig(22), gate(ASub, 21, 6, 22), conn(22, 2)]

CASES = [("callee_before_wiring", P1_SRC, P1_EVENTS), ("nested_component_array", P2_SRC, P2_EVENTS), ("lessthan_bits", P3_SRC, P3_EVENTS)]


def _arr(ev):
    return np.asarray(ev, dtype=np.uint32).reshape(-1, 4)


@pytest.mark.parametrize("name,src,events", CASES, ids=[c[0] for c in CASES])
def test_front_end_records_the_hand_derived_calls(c2a, name, src, events):
    import ctypes as C_
    lib = c2a.lib
    p = lib.c2a_program_new()
    try:
        assert lib.c2a_program_compile_source(p, src.encode(), None, None) == 0, lib.c2a_program_error(p)
        n = lib.c2a_program_num_events(p)
        got = np.ctypeslib.as_array(C_.cast(lib.c2a_program_events(p), C_.POINTER(C_.c_uint32)), shape=(n, 4)).copy()
    finally:
        lib.c2a_program_free(p)
    want = _arr(events)
    assert got.shape == want.shape, (got.tolist(), want.tolist())
    assert np.array_equal(got, want), f"first difference at call {int(np.argmax((got != want).any(axis=1)))}: {got.tolist()} vs {want.tolist()}"


@pytest.mark.parametrize("name,src,events", CASES, ids=[c[0] for c in CASES])
def test_oracle_and_host_emitter_agree_on_the_hand_derived_calls(c2a, orc, name, src, events):
    ev = _arr(events)
    oc = orc.OracleCompiler()
    oc.emit_events(ev)
    comp = c2a.Compiler.__new__(c2a.Compiler)   # host emitter only: no device context needed
    comp._c = c2a.lib.c2a_compiler_new()
    comp._ctx = None
    try:
        import ctypes as C_
        bad = C_.c_uint64(0)
        assert c2a.lib.c2a_emit_events(comp._c, ev.ctypes.data_as(C_.c_void_p), ev.shape[0], C_.byref(bad)) == 0
        Gn = c2a.lib.c2a_num_gates(comp._c)
        hg = np.empty((Gn, 4), dtype=np.uint32)
        c2a.lib.c2a_get_gates(comp._c, hg.ctypes.data_as(C_.c_void_p))
        assert np.array_equal(hg, oc.gate_array()) and c2a.lib.c2a_node_count(comp._c) == oc.node_count
    finally:
        c2a.lib.c2a_compiler_free(comp._c)
        comp._c = None


def test_first_program_by_hand_down_to_the_built_circuit(orc):
    """node ids, DFS order, first-seen wires and renumbered gates of P1, all derived by hand (module docstring)"""
    oc = orc.OracleCompiler()
    oc.emit_events(_arr(P1_EVENTS))
    assert oc.gate_array().tolist() == P1_GATES and oc.node_count == P1_NODE_COUNT
    st, _, order, wire, ng, wc = orc.backend_raw(oc.gate_array(), P1_NODE_COUNT + 1, [oc.signal_node(0)], [oc.signal_node(1)])
    assert st == 0 and order.tolist() == P1_ORDER and ng.tolist() == P1_NEW_GATES and wc == P1_WIRE_COUNT
    assert [int(wire[oc.signal_node(s)]) for s in (0, 5, 6, 1)] == [0, 1, 2, 3]


@pytest.mark.gpu
@pytest.mark.parametrize("name,src,events", CASES, ids=[c[0] for c in CASES])
def test_device_paths_on_the_hand_derived_calls(c2a, ctx, orc, name, src, events):
    """device emitter (16-byte events, both packed forms, the fused kernel) + build against the oracle on the hand-derived calls;
    and the whole product path from the .circom text"""
    ev = _arr(events)
    oc = orc.OracleCompiler()
    oc.emit_events(ev)
    kinds = ev[:, 0] & 0xFF
    io_in = [0, 1] if name != "callee_before_wiring" else [0]
    io_out = {"callee_before_wiring": [1], "nested_component_array": [2, 3], "lessthan_bits": [2]}[name]
    nodes = lambda s: np.array([oc.signal_node(int(x)) for x in s], dtype=np.uint32)
    st, _, o_order, o_wire, o_gates, o_wc = orc.backend_raw(oc.gate_array(), oc.node_count + 1, nodes(io_in), nodes(io_out))
    assert st == 0
    info = ctx.emit_events(ev)
    g, _nos = ctx.emitted_fetch()
    assert info["path"] == 1 and np.array_equal(g, oc.gate_array()) and info["node_count"] == oc.node_count
    for implicit in (False, True):
        k, w, f = c2a.pack_events(ev, implicit=implicit)
        for fused in (0, 1 << 22):
            c2a.lib.c2a_set_fused_limits(fused, 0)
            try:
                _i, order, wire, ng, wc = ctx.compile_packed(k, w, f, io_in, io_out)
            finally:
                c2a.lib.c2a_set_fused_limits(1 << 22, 0)
            assert np.array_equal(order, o_order) and np.array_equal(wire, o_wire) and np.array_equal(ng, o_gates) and wc == o_wc
    if name == "callee_before_wiring":
        assert order.tolist() == P1_ORDER and ng.tolist() == P1_NEW_GATES and wc == P1_WIRE_COUNT
    dev = c2a.compile(None, source=src, context=ctx, emitter="device")
    circ = dev.build_circuit()
    assert np.array_equal(dev.events, ev) and np.array_equal(circ.gate_array, o_gates) and circ.wire_count == o_wc
    del kinds
