"""The product's union-find emitter (csrc/c2a_host.cpp) against the faithful linear-scan oracle on random
event streams: gates, node ids, per-node signal order, flags and error behaviour must be identical
(reference: src/compiler.rs:139-278).  No GPU involved."""
import numpy as np
import pytest


def random_stream(rng, n_events, n_ids, p_unknown=0.03):
    """list of ('S', id, value|None) | ('G', op, l, r, o) | ('C', a, b); includes duplicates / unknown ids."""
    ev = []
    declared = []
    for _ in range(n_events):
        x = rng.rand()
        if x < 0.35 or len(declared) < 3:
            i = int(rng.randint(0, n_ids))
            ev.append(("S", i, int(rng.randint(0, 50)) if rng.rand() < 0.15 else None))
            declared.append(i)
        elif x < 0.65:
            pick = lambda: int(rng.randint(0, n_ids + 5)) if rng.rand() < p_unknown else int(declared[rng.randint(len(declared))])
            ev.append(("G", int(rng.randint(0, 20)), pick(), pick(), pick()))
        else:
            pick = lambda: int(rng.randint(0, n_ids + 5)) if rng.rand() < p_unknown else int(declared[rng.randint(len(declared))])
            ev.append(("C", pick(), pick()))
    return ev


def replay(C, E, ev):
    c = C()
    log = []
    for e in ev:
        try:
            if e[0] == "S":
                c.add_signal(e[1], f"s{e[1]}", e[2])
            elif e[0] == "G":
                c.add_gate(e[1], e[2], e[3], e[4])
            else:
                c.add_connection(e[1], e[2])
            log.append(0)
        except E as ex:
            log.append(int(ex.status))
    return c, log


@pytest.mark.parametrize("seed", range(40))
def test_random_streams_match_oracle(c2a, orc, seed):
    rng = np.random.RandomState(seed)
    ev = random_stream(rng, n_events=int(rng.randint(20, 400)), n_ids=int(rng.randint(5, 120)), p_unknown=0.05 if seed % 3 == 0 else 0.0)
    a, la = replay(orc.OracleCompiler, orc.OracleError, ev)
    b, lb = replay(c2a.Compiler, c2a.CircuitError, ev)
    assert la == lb
    assert a.node_count == b.node_count
    assert a.gate_array().tolist() == b.gate_array().tolist()
    assert a.nodes() == b.nodes()
    for sid in range(0, 130):
        assert a.signal_node(sid) == b.signal_node(sid)


def test_bulk_events_match_single_calls(c2a, orc):
    wl = c2a.workloads.mimc_chains(3, rounds=5, variant="late")
    a = orc.OracleCompiler()
    a.emit_events(wl.events)
    b = c2a.Compiler()
    b.emit_events(wl.events)
    assert a.gate_array().tolist() == b.gate_array().tolist()
    assert a.nodes() == b.nodes()
    assert b.gate_array().shape[0] == wl.n_gates == 3 * (6 * 5 + 1)
    assert b.num_signals == wl.n_signals


def test_sparse_signal_ids(c2a, orc):
    ids = [0, 5, 4_000_000_000, 123_456_789, 7]
    for C in (orc.OracleCompiler, c2a.Compiler):
        c = C()
        for i in ids:
            c.add_signal(i, f"s{i}", None)
        c.add_gate(0, ids[2], ids[3], ids[4])
        c.add_connection(ids[2], ids[0])
        assert c.gate_array().tolist() == [[0, 6, 4, 5]]


def test_get_signals_prefix_match(c2a):  # compiler.rs:676-690
    c = c2a.Compiler()
    c.add_signal(1, "signal1", None)
    c.add_signal(2, "filter_signal", None)
    assert c.get_signals("filter") == {2: "filter_signal"}


def test_bristol_gate_lines_are_formatted_natively(c2a):
    """circuit.txt body (src/main.rs:34-36 -> bristol-circuit write_bristol; layout parity unpinned): the native formatter against
    the obvious Python formatting, including 1- and 10-digit wire ids and every op token"""
    import numpy as np
    from circom_2_arithc_b200 import program as P
    from circom_2_arithc_b200.compiler import BristolCircuit, CircuitInfo
    rng = np.random.RandomState(3)
    G = 5000
    g = np.zeros((G, 4), dtype=np.uint32)
    g[:, 0] = np.arange(G) % 20
    g[:, 1:] = rng.randint(0, 2 ** 32, size=(G, 3), dtype=np.int64).astype(np.uint32)
    g[:4, 1:] = [[0, 1, 9], [10, 99, 100], [4294967295, 0, 7], [1000000000, 999999999, 123456789]]
    ci = CircuitInfo()
    ci.input_name_to_wire_index = {"0.a": 0, "0.b": 1}
    ci.output_name_to_wire_index = {"0.c": 2}
    names = [t.name for t in c2a.AGateType]
    want = "\n".join([f"{G} 77", "2 1 1", "1 1", ""] + [f"2 1 {a} {b} {o} {names[op]}" for op, a, b, o in g.tolist()]) + "\n"
    assert P.bristol_text(BristolCircuit(wire_count=77, info=ci, gate_array=g, order=np.arange(G, dtype=np.uint32))) == want
    empty = BristolCircuit(wire_count=0, info=CircuitInfo(), gate_array=np.zeros((0, 4), np.uint32), order=np.zeros(0, np.uint32))
    assert P.bristol_text(empty) == "0 0\n0\n0\n\n"
