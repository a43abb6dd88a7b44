"""Worst-case shapes of the exact-order sort (src/topological_sort.rs:23-50) at 1 M gates, through the C ABI, bit-exact against the
oracle AND not slower than the oracle's own (recursion-free) DFS on the same input: one DFS tree holding every gate (a reversed
chain), a forward chain through the rh operand, a single forward edge on top of a 1 M-deep in-order chain, a big fan-in tree emitted
root first, a tree block with a cycle behind it, and a big block that is a DAG (shared gates: the one-thread DFS path).

What makes them hard on a GPU: r[] (the DFS root that first reaches a gate) is a minimum over all transitive consumers - a label
that has to travel down a 1 M-long dependency chain - and the post-order of a 1 M-gate block is one sequential walk.  The product
path (csrc/c2a_device.cu) bounds every data-driven walk (k_relax_loop: rounds on device-resident queues, a per-thread stack for the rh side, pointer jumping along the
smallest-index consumer when the queues do not drain) and emits tree-shaped big blocks by pointer jumping (k_tree_blocks)."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

A = 0  # AAdd


def _check(ctx, orc, c2a, gates, nb, ins, outs, faster_than_oracle=True, expect_robust=None):
    gates = np.ascontiguousarray(gates, dtype=np.uint32)
    t0 = time.perf_counter()
    st, err, o_order, o_wire, o_gates, o_wc = orc.backend_raw(gates, nb, ins, outs)
    t_orc = time.perf_counter() - t0
    if st != 0:
        with pytest.raises(c2a.CircuitError) as e:
            ctx.build_circuit(gates, nb, ins, outs)
        assert e.value.message == f"detected at i={err}"
        return
    ctx.build_circuit(gates, nb, ins, outs)   # warm-up: the first call grows the scratch slab
    t0 = time.perf_counter()
    order, wire, ng, wc = ctx.build_circuit(gates, nb, ins, outs)
    t_gpu = time.perf_counter() - t0          # host buffers: H2D / D2H of every array included
    ph = ctx.phases()
    assert np.array_equal(order, o_order), f"order differs first at {int(np.argmax(order != o_order))}"
    assert wc == o_wc and np.array_equal(wire, o_wire) and np.array_equal(ng, o_gates)
    if expect_robust is not None:
        assert bool(int(ph["n_relax_fallback_rounds"]) & 0x10000) == expect_robust, ph
    if faster_than_oracle:
        assert t_gpu <= t_orc, f"GPU call {t_gpu * 1e3:.1f} ms (phases {ph}) vs oracle {t_orc * 1e3:.1f} ms"
    return t_gpu, t_orc, ph


def _chain(G, slot, reverse=True):
    """gate g writes node 10+g and reads the node of gate g+1 (reverse) / g-1 through operand `slot`; the other operand is input 1"""
    gates = np.zeros((G, 4), dtype=np.uint32)
    gates[:, 0] = 7
    gates[:, 3] = 10 + np.arange(G)
    other = 3 - slot
    gates[:, other] = 1
    if reverse:
        gates[:, slot] = 10 + np.arange(G) + 1
        gates[G - 1, slot] = 2
    else:
        gates[:, slot] = 10 + np.arange(G) - 1
        gates[0, slot] = 2
    return gates


@pytest.mark.parametrize("slot", [1, 2])
def test_reversed_chain_one_million_gates(ctx, orc, c2a, slot):
    """gate 0 is the LAST link: root 0 reaches every gate - one DFS tree of 1 M gates, 1 M forward edges (lh or rh operand)"""
    G = 1_000_000
    res = _check(ctx, orc, c2a, _chain(G, slot), 10 + G + 1, [1, 2], [10], expect_robust=True)
    assert res is not None


def test_one_forward_edge_on_top_of_a_deep_in_order_chain(ctx, orc, c2a):
    """gates 1..G-1 form an in-order chain (no forward edge); gate 0 reads the END of it: a single seed whose label has to travel
    down 1 M dependency hops"""
    G = 1_000_000
    gates = _chain(G, 1, reverse=False)
    gates[1, 1] = 2                      # the chain starts at gate 1 ...
    gates[0, 1] = 10 + G - 1             # ... and gate 0 consumes its last gate
    _check(ctx, orc, c2a, gates, 10 + G + 1, [1, 2], [10], expect_robust=True)


def test_big_fan_in_tree_emitted_root_first(ctx, orc, c2a):
    """a complete binary tree of 2^19 - 1 gates in heap order (gate g reads gates 2g+1 and 2g+2): every edge points forward, one
    block, depth 19 - branching exercises the lh-before-rh offsets of the closed-form post-order"""
    G = (1 << 19) - 1
    g = np.arange(G)
    gates = np.zeros((G, 4), dtype=np.uint32)
    gates[:, 0] = A
    gates[:, 3] = 10 + g
    l, r = 2 * g + 1, 2 * g + 2
    gates[:, 1] = np.where(l < G, 10 + l, 1)
    gates[:, 2] = np.where(r < G, 10 + r, 2)
    _check(ctx, orc, c2a, gates, 10 + G, [1, 2], [10])
    # the same tree with the two operands swapped on every other level, and a few gates reading one child twice
    gates2 = gates.copy()
    lvl = np.floor(np.log2(g + 1)).astype(np.int64)
    sw = (lvl % 2 == 1)
    gates2[sw, 1], gates2[sw, 2] = gates[sw, 2], gates[sw, 1]
    dbl = (g % 1001 == 5) & (l < G)
    gates2[dbl, 2] = gates2[dbl, 1]      # rh == lh: that child is visited once, the other subtree becomes blocks of its own
    _check(ctx, orc, c2a, gates2, 10 + G, [1, 2], [10], faster_than_oracle=False)


def test_caterpillar_two_children_per_level(ctx, orc, c2a):
    """a 300 K-deep spine whose every gate also reads a private leaf gate through the other operand (alternating sides): a tree of
    depth 300 K where lh / rh order matters at every level"""
    D = 300_000
    G = 2 * D
    gates = np.zeros((G, 4), dtype=np.uint32)
    gates[:, 0] = A
    gates[:, 3] = 10 + np.arange(G)
    spine = np.arange(D)                  # gates 0..D-1: spine (gate s reads spine s+1 and leaf D+s), gates D..2D-1: leaves
    side = spine % 2
    nxt = np.where(spine + 1 < D, 10 + spine + 1, 1)
    leaf = 10 + D + spine
    gates[spine, 1] = np.where(side == 0, nxt, leaf)
    gates[spine, 2] = np.where(side == 0, leaf, nxt)
    gates[D:, 1] = 1
    gates[D:, 2] = 2
    _check(ctx, orc, c2a, gates, 10 + G, [1, 2], [10])


def test_cycle_behind_a_big_tree_block(ctx, orc, c2a):
    """a 200 K reversed chain whose far end closes a cycle: the block is no tree (its root gets an in-block consumer), the one-thread
    DFS must find the reference's `detected at i=`"""
    G = 200_000
    gates = _chain(G, 1)
    gates[G - 1, 1] = 10 + 5              # the last gate reads gate 5: a cycle 5 -> 6 -> ... -> G-1 -> 5
    _check(ctx, orc, c2a, gates, 10 + G + 1, [1, 2], [10])


def test_big_dag_block_with_shared_gates(ctx, orc, c2a):
    """a 100 K-deep ladder emitted top first: gate g reads gates g+1 and g+2 - every gate has two consumers in the block, so it is
    no tree and goes through the sequential DFS (correct, not fast: no timing claim)"""
    G = 100_000
    gates = np.zeros((G, 4), dtype=np.uint32)
    gates[:, 0] = A
    gates[:, 3] = 10 + np.arange(G)
    gates[:, 1] = np.where(np.arange(G) + 1 < G, 10 + np.arange(G) + 1, 1)
    gates[:, 2] = np.where(np.arange(G) + 2 < G, 10 + np.arange(G) + 2, 2)
    _check(ctx, orc, c2a, gates, 10 + G, [1, 2], [10], faster_than_oracle=False)


def test_shuffled_chains_many_relaxation_rounds(ctx, orc, c2a):
    """BASELINE config 5's stress variant at 1 M gates: the gate vector of 1 832 MiMC chains shuffled - half the gates hold a forward
    edge, the relaxation needs dozens of rounds (device-resident queue counters, no host round trip)"""
    wl = c2a.workloads.mimc_chains(1832, rounds=91, variant="late")
    comp = c2a.Compiler(context=ctx)
    comp.emit_events(wl.events)
    gates = c2a.workloads.shuffle_gates(comp.gate_array(), seed=1)
    ins = comp.signal_nodes(np.array(sorted(wl.inputs), dtype=np.uint32))
    outs = comp.signal_nodes(np.array(sorted(wl.outputs), dtype=np.uint32))
    _check(ctx, orc, c2a, gates, comp.node_count + 1, ins, outs)
