"""Pins the CPU oracle (oracle/c2a_oracle.cpp) against the reference's OWN tests, restated:
  src/compiler.rs:584-795        unit tests of Node / add_signal / add_gate / add_connection
  tests/integration.rs:279-441   integration tests (exact constants / output maps, functional simulations)
  src/process.rs:772-822         execute_op known answers
plus the hand-derived goldens of SURVEY.md §4.  The same assertions are run against the product's host emitter
(no GPU needed for the emit side)."""
import numpy as np
import pytest

import miniwalker as mw


def both(c2a, orc):
    return [("oracle", orc.OracleCompiler, orc.OracleError), ("product", c2a.Compiler, c2a.CircuitError)]


@pytest.fixture(params=["oracle", "product"])
def impl(request, c2a, orc):
    return {"oracle": (orc.OracleCompiler, orc.OracleError), "product": (c2a.Compiler, c2a.CircuitError)}[request.param]


# ---- src/compiler.rs unit tests -------------------------------------------------------------------------
def test_compiler_add_signal(impl):  # compiler.rs:650-664
    C, _ = impl
    c = C()
    c.add_signal(1, "signal1", None)
    assert c.num_signals == 1
    assert c.nodes() == {1: {"is_const": False, "is_out": False, "signals": [1]}}


def test_compiler_add_duplicated_signal(impl):  # compiler.rs:666-674
    C, E = impl
    c = C()
    c.add_signal(1, "signal1", None)
    with pytest.raises(E) as e:
        c.add_signal(1, "signal1", None)
    assert e.value.status == 3  # SignalAlreadyDeclared


def test_compiler_add_gate(impl):  # compiler.rs:692-713: gates record NODE ids 1,2,3
    C, _ = impl
    c = C()
    for i in (1, 2, 3):
        c.add_signal(i, f"signal{i}", None)
    c.add_gate(mw.AAdd, 1, 2, 3)
    assert c.gate_array().tolist() == [[mw.AAdd, 1, 2, 3]]
    assert c.nodes()[3]["is_out"] is True  # compiler.rs:201


def test_compiler_add_connection(impl):  # compiler.rs:716-739: merged node gets id 4 and holds both signals
    C, _ = impl
    c = C()
    for i in (1, 2, 3):
        c.add_signal(i, f"signal{i}", None)
    c.add_connection(1, 2)
    n = c.nodes()
    assert len(n) == 2 and n[4]["signals"] == [1, 2]


def test_compiler_add_connection_same_node(impl):  # compiler.rs:742-758
    C, _ = impl
    c = C()
    c.add_signal(1, "signal1", None)
    c.add_signal(2, "signal2", None)
    c.add_connection(1, 2)
    c.add_connection(1, 2)
    assert len(c.nodes()) == 1 and c.node_count == 3


def test_compiler_add_connection_output_nodes(impl):  # compiler.rs:761-777
    C, E = impl
    c = C()
    for i in (1, 2, 3, 4):
        c.add_signal(i, f"s{i}", None)
    c.add_gate(mw.AAdd, 3, 4, 1)  # marks node(1) as out
    c.add_gate(mw.AAdd, 3, 4, 2)  # marks node(2) as out
    with pytest.raises(E) as e:
        c.add_connection(1, 2)
    assert e.value.status == 4  # CannotMergeOutputNodes


def test_compiler_add_connection_constant_nodes(impl):  # compiler.rs:780-794
    C, E = impl
    c = C()
    c.add_signal(1, "signal1", 1)
    c.add_signal(2, "signal2", 2)
    with pytest.raises(E) as e:
        c.add_connection(1, 2)
    assert e.value.status == 5  # CannotMergeConstantNodes


def test_gate_rewrite_after_merge(impl):  # compiler.rs:260-270: gates follow merged node ids
    C, _ = impl
    c = C()
    for i in range(5):
        c.add_signal(i, f"s{i}", None)          # nodes 1..5
    c.add_gate(mw.AMul, 0, 1, 2)               # (AMul,1,2,3)
    c.add_connection(2, 3)                      # node 6 = [2,3]
    c.add_connection(0, 4)                      # node 7 = [0,4]
    assert c.gate_array().tolist() == [[mw.AMul, 7, 2, 6]]
    assert c.nodes()[6] == {"is_const": False, "is_out": True, "signals": [2, 3]}


def test_unknown_signals_resolve_to_node_zero(impl):  # compiler.rs:183 (nodes default to id 0), :201 unwrap
    C, E = impl
    c = C()
    c.add_signal(0, "a", None)
    c.add_signal(1, "b", None)
    c.add_gate(mw.AAdd, 77, 0, 1)          # unknown lhs -> node id 0
    assert c.gate_array().tolist() == [[mw.AAdd, 0, 1, 2]]
    with pytest.raises(E) as e:
        c.add_gate(mw.AAdd, 0, 1, 99)      # unknown out: the reference panics
    assert e.value.status == 6
    c.add_connection(55, 0)                # merging "node 0" with node(a): gates that held 0 follow (compiler.rs:260-270)
    assert c.gate_array().tolist() == [[mw.AAdd, 3, 3, 2]]
    c.add_gate(mw.AAdd, 78, 0, 1)          # a later unknown is node 0 again
    assert c.gate_array().tolist()[1] == [mw.AAdd, 0, 3, 2]
    c.add_connection(55, 56)               # both unknown: 0 == 0, nothing happens
    assert c.node_count == 3


# ---- SURVEY.md §4 hand goldens (emit side) -----------------------------------------------------------------
def test_golden_add_zero_emit(impl):
    C, _ = impl
    c = C()
    mw.fixture_add_zero(c)
    assert c.gate_array().tolist() == [[mw.AAdd, 1, 3, 5]] and c.node_count == 5
    assert c.nodes()[5]["signals"] == [3, 1]


def test_golden_sum_emit(impl):
    C, _ = impl
    c = C()
    mw.fixture_sum(c)
    assert c.gate_array().tolist() == [[mw.AAdd, 1, 2, 5]] and c.node_count == 5


def test_golden_constant_sum_emit(impl):
    C, _ = impl
    c = C()
    mw.fixture_constant_sum(c)
    assert c.gate_array().shape[0] == 0 and c.node_count == 3
    assert c.nodes() == {3: {"is_const": True, "is_out": False, "signals": [1, 0]}}


def test_golden_array_assignment_emit(impl):
    C, _ = impl
    c = C()
    mw.fixture_array_assignment(c)
    assert c.gate_array().tolist() == [[mw.AAdd, 15, 16, 11], [mw.AAdd, 11, 17, 12], [mw.AAdd, 12, 18, 19]]
    assert c.node_count == 19
    assert c.nodes()[19]["signals"] == [12, 9, 4]


# ---- oracle back end: tests/integration.rs -----------------------------------------------------------------
def run_named(orc, circ, inputs):
    info = circ["info"]
    vals = {info["input_name_to_wire_index"][k]: v for k, v in inputs.items()}
    for _, ci in info["constants"].items():
        vals[ci["wire_index"]] = int(ci["value"])
    wires = orc.simulate(circ["gates"], circ["wire_count"], vals)
    return {k: wires[w] for k, w in info["output_name_to_wire_index"].items()}


def test_integration_add_zero(orc):  # integration.rs:279-286
    c = orc.OracleCompiler()
    mw.fixture_add_zero(c)
    circ = c.build_circuit()
    assert circ["wire_count"] == 3 and circ["gates"].tolist() == [[mw.AAdd, 0, 1, 2]]
    assert circ["info"]["constants"] == {"0.const_signal_0_2": {"value": "0", "wire_index": 1}}
    assert run_named(orc, circ, {"0.in": 42}) == {"0.out": 42}


def test_integration_infix_ops(orc):  # integration.rs:288-333
    c = orc.OracleCompiler()
    mw.fixture_infix_ops(c)
    circ = c.build_circuit()
    out = run_named(orc, circ, {f"0.x{i}": i for i in range(6)})
    assert out == {f"0.{n}": e for n, _o, _l, _r, e in mw.INFIX_OUTPUTS}
    assert circ["order"].tolist() == list(range(29))


def test_integration_matrix_element_multiplication(orc):  # integration.rs:335-356
    c = orc.OracleCompiler()
    mw.fixture_mat_elem_mul(c)
    circ = c.build_circuit()
    ins = {f"0.{m}[{i}][{j}]": 2 for m in "ab" for i in range(2) for j in range(2)}
    assert run_named(orc, circ, ins) == {f"0.out[{i}][{j}]": 4 for i in range(2) for j in range(2)}


def test_integration_sum(orc):  # integration.rs:358-365
    c = orc.OracleCompiler()
    mw.fixture_sum(c)
    circ = c.build_circuit()
    assert circ["wire_count"] == 3
    assert run_named(orc, circ, {"0.a": 3, "0.b": 5}) == {"0.out": 8}


def test_integration_x_eq_x(orc):  # integration.rs:367-374
    c = orc.OracleCompiler()
    mw.fixture_x_eq_x(c)
    assert run_named(orc, c.build_circuit(), {"0.x": 37}) == {"0.out": 1}


def test_integration_constant_sum(orc):  # integration.rs:393-415 (exact)
    c = orc.OracleCompiler()
    mw.fixture_constant_sum(c)
    circ = c.build_circuit()
    assert circ["info"]["constants"] == {"0.const_signal_8_1": {"value": "8", "wire_index": 0}}


def test_integration_direct_output(orc):  # integration.rs:417-441 (exact)
    c = orc.OracleCompiler()
    mw.fixture_direct_output(c)
    circ = c.build_circuit()
    assert circ["info"]["output_name_to_wire_index"] == {"0.out": 0}
    assert circ["info"]["constants"] == {"0.const_signal_42_1": {"value": "42", "wire_index": 0}}


def test_integration_prefix_ops_negative_golden(orc):  # integration.rs:455-462: prefix-match I/O tagging clash
    c = orc.OracleCompiler()
    mw.fixture_prefix_ops(c)
    with pytest.raises(orc.OracleError) as e:
        c.build_circuit()
    assert e.value.status == 2
    assert "used for both input 0.complement" in e.value.message and "and output 0.complement" in e.value.message


def test_golden_array_assignment_build(orc):  # SURVEY.md §4 table, last row
    c = orc.OracleCompiler()
    mw.fixture_array_assignment(c)
    circ = c.build_circuit()
    assert circ["order"].tolist() == [0, 1, 2] and circ["wire_count"] == 7
    assert circ["gates"].tolist() == [[mw.AAdd, 0, 1, 4], [mw.AAdd, 4, 2, 5], [mw.AAdd, 5, 3, 6]]
    assert run_named(orc, circ, {f"0.a_in[{i}][{j}]": 1 + 2 * i + j for i in range(2) for j in range(2)}) == {"0.out": 10}


def test_report_matches_reference_rules(orc):  # compiler.rs:287-319, 503-531
    c = orc.OracleCompiler()
    mw.fixture_add_zero(c)
    rep = c.report()
    assert rep["value_type"] == "sint"
    assert [r["id"] for r in rep["inputs"]] == [1, 3] and rep["inputs"][1]["value"] == 0
    assert rep["outputs"] == [{"id": 5, "names": ["0.out"], "value": None}]  # random_ names filtered (:519)


# ---- topological_sort.rs ---------------------------------------------------------------------------------
def test_topological_sort_order(orc):
    # roots ascending, deps in the given order, post-order push (topological_sort.rs:11-13, 42-47)
    assert orc.topological_sort([[2, 1], [3], [3], []]).tolist() == [3, 2, 1, 0]
    assert orc.topological_sort([[], [0], [1, 0]]).tolist() == [0, 1, 2]
    assert orc.topological_sort([[1], [2], []]).tolist() == [2, 1, 0]


def test_topological_sort_cycle(orc):  # topological_sort.rs:34-38
    with pytest.raises(orc.OracleError) as e:
        orc.topological_sort([[1], [2], [1]])
    assert e.value.status == 1 and e.value.message == "detected at i=1"
    with pytest.raises(orc.OracleError) as e:
        orc.topological_sort([[], [1]])
    assert e.value.message == "detected at i=1"


# ---- src/process.rs:772-822 execute_op -------------------------------------------------------------------
def test_execute_op_kats(orc):
    kats = [(3, 4, mw.AAdd, 7), (10, 5, mw.ASub, 5), (6, 3, mw.AMul, 18), (9, 3, mw.ADiv, 3), (7, 3, mw.AMod, 1), (2, 3, mw.APow, 8),
            (8, 2, mw.AShiftL, 32), (8, 2, mw.AShiftR, 2), (5, 5, mw.AEq, 1), (5, 4, mw.ANeq, 1), (1, 0, mw.ABoolOr, 1),
            (1, 1, mw.ABoolAnd, 1), (1, 1, mw.ABitOr, 1), (1, 1, mw.ABitAnd, 1), (1, 1, mw.AXor, 0)]
    for l, r, op, want in kats:
        assert orc.execute_op(l, r, op)[:2] == (0, want)
    for op in (mw.ADiv, mw.AIntDiv, mw.AMod):
        assert orc.execute_op(10, 0, op)[0] == 1
    assert orc.execute_op(0, 5, mw.ASub) == (1, 0, "Subtraction underflow")[:1] + orc.execute_op(0, 5, mw.ASub)[1:]
    assert orc.execute_op(0, 5, mw.ASub)[2] == "Subtraction underflow"
    assert orc.execute_op(0, 0, mw.AEq)[1] == 1 and orc.execute_op(0, 1, mw.AEq)[1] == 0            # !0, !1
    assert orc.execute_op(0xFFFFFFFF, 0b1010, mw.AXor)[1] == 0b1111_1111_1111_1111_1111_1111_1111_0101  # ~0b1010
