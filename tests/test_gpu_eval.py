"""GPU evaluator (c2a_evaluate, csrc/c2a_eval.cuh) against the oracle's straight-line restatement of the reference's
test-side simulator (tests/integration.rs:90-119, 191-237): same wire values, same failure position."""
import numpy as np
import pytest

import miniwalker as mw

pytestmark = pytest.mark.gpu


def both(ctx, orc, c2a, gates, wire_count, values):
    """run both; returns the GPU wire map (or the failing gate position)"""
    try:
        want = orc.simulate(gates, wire_count, values)
        fail_at = None
    except orc.OracleError as e:
        fail_at = e.status - 1
    if fail_at is None:
        got = ctx.evaluate(gates, wire_count, values)
        assert got == want
        return got
    with pytest.raises(c2a.C2AError) as ex:
        ctx.evaluate(gates, wire_count, values)
    assert ex.value.status == c2a.Status.EVALUATION and ex.value.err_index == fail_at
    return fail_at


def test_reference_simulations_evaluated_on_gpu(ctx, orc, c2a):
    """tests/integration.rs:279-374 end to end on the device: compile -> build_circuit -> evaluate"""
    def run(fx, inputs):
        c = c2a.Compiler(context=ctx)
        fx(c)
        circ = c.build_circuit()
        vals = {circ.info.input_name_to_wire_index[k]: v for k, v in inputs.items()}
        for ci in circ.info.constants.values():
            vals[ci.wire_index] = int(ci.value)
        wires = both(ctx, orc, c2a, circ.gate_array, circ.wire_count, vals)
        return {k: wires[w] for k, w in circ.info.output_name_to_wire_index.items()}
    assert run(mw.fixture_add_zero, {"0.in": 42}) == {"0.out": 42}
    assert run(mw.fixture_infix_ops, {f"0.x{i}": i for i in range(6)}) == {f"0.{n}": e for n, _o, _l, _r, e in mw.INFIX_OUTPUTS}
    assert run(mw.fixture_sum, {"0.a": 3, "0.b": 5}) == {"0.out": 8}
    assert run(mw.fixture_x_eq_x, {"0.x": 37}) == {"0.out": 1}
    ins = {f"0.{m}[{i}][{j}]": 2 for m in "ab" for i in range(2) for j in range(2)}
    assert run(mw.fixture_mat_elem_mul, ins) == {f"0.out[{i}][{j}]": 4 for i in range(2) for j in range(2)}


def random_ssa(rng, G, n_in, p_bad=0.0, ops=None, window=0):
    gates = np.zeros((G, 4), dtype=np.uint32)
    W = n_in + G + 3  # 3 wires that nobody sets
    for g in range(G):
        lo = max(0, n_in + g - window) if window else 0
        a, b = (int(rng.randint(lo, n_in + g)) for _ in range(2))
        if rng.rand() < p_bad:
            a = int(rng.choice([W - 1, W - 2, n_in + g, min(n_in + g + 5, n_in + G - 1)]))  # unset / itself / written later
        gates[g] = (int(rng.choice(ops)) if ops is not None else int(rng.randint(0, 20)), a, b, n_in + g)
    return gates, W


@pytest.mark.parametrize("seed", range(20))
def test_random_single_assignment_circuits(ctx, orc, c2a, seed):
    rng = np.random.RandomState(seed)
    G = int(rng.choice([1, 7, 100, 3000, 40000]))
    n_in = int(rng.randint(1, 20))
    safe = [0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 14, 15, 16, 17, 18, 19]  # no div / mod
    mode = seed % 4
    gates, W = random_ssa(rng, G, n_in, p_bad=0.002 if mode == 3 else 0.0, ops=safe if mode in (0, 3) else None, window=64 if seed % 2 else 0)
    values = {i: int(rng.randint(0, 2 ** 32)) if rng.rand() < 0.7 else int(rng.randint(0, 3)) for i in range(n_in)}
    res = both(ctx, orc, c2a, gates, W, values)
    if mode == 0:
        assert isinstance(res, dict) and len(res) == n_in + G


def test_not_single_assignment_is_rejected(ctx, c2a):
    g = np.array([[0, 0, 1, 2], [0, 0, 1, 2]], dtype=np.uint32)
    with pytest.raises(c2a.C2AError) as ex:
        ctx.evaluate(g, 3, {0: 1, 1: 2})
    assert ex.value.status == c2a.Status.INVALID_ARGUMENT
    g = np.array([[0, 0, 1, 1]], dtype=np.uint32)  # overwrites an input
    with pytest.raises(c2a.C2AError):
        ctx.evaluate(g, 2, {0: 1, 1: 2})
    assert ctx.evaluate(np.zeros((0, 4), np.uint32), 2, {1: 9}) == {1: 9}


def _emit_build_eval(ctx, orc, c2a, wl, seed=0):
    rng = np.random.RandomState(seed)
    info = ctx.emit_events(np.ascontiguousarray(wl.events))
    ins = np.array(sorted(wl.inputs), dtype=np.uint32)
    outs = np.array(sorted(wl.outputs), dtype=np.uint32)
    _, nos = ctx.emitted_fetch(want_gates=False)
    order, wire, ng, wc = ctx.emitted_build_circuit(ins, outs)
    values = {int(wire[nos[s]]): int(rng.randint(0, 2 ** 32)) for s in ins}
    ev = wl.events
    consts = ev[(ev[:, 0] & 0xFF) == 1]
    for sid, v in zip(consts[:, 1], consts[:, 2]):
        w = int(wire[nos[sid]])
        if w != 0xFFFFFFFF:
            values[w] = int(v)
    return both(ctx, orc, c2a, ng, wc, values), wire, nos, outs


def test_deep_narrow_circuit_uses_the_single_cta_levels(ctx, orc, c2a):
    res, wire, nos, outs = _emit_build_eval(ctx, orc, c2a, c2a.workloads.sha256_shaped(rounds=8))
    assert isinstance(res, dict) and all(int(wire[nos[s]]) in res for s in outs)
    assert "k_eval_level" in ctx.phases()


def test_mimc_one_million_gates(ctx, orc, c2a):
    res, wire, nos, outs = _emit_build_eval(ctx, orc, c2a, c2a.workloads.mimc_chains(1832, rounds=91, variant="late"), seed=3)
    assert isinstance(res, dict) and all(int(wire[nos[s]]) in res for s in outs)
