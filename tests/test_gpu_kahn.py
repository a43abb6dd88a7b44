"""K3/K4 (consumer CSR + Kahn frontier levels) and the opt-in K8/K9 layer-wise sweeps, against small CPU
restatements written here (these are not reference features: the reference has no Kahn sort and no sweeps;
the dependency relation is the reference's, src/compiler.rs:401-421)."""
import numpy as np
import pytest

import miniwalker as mw

pytestmark = pytest.mark.gpu


def cpu_levels(orc, gates, node_bound):
    gates = np.asarray(gates, dtype=np.uint32).reshape(-1, 4)
    G = gates.shape[0]
    prod = np.full(node_bound, -1, dtype=np.int64)
    prod[gates[:, 3]] = np.arange(G)
    st, err, order, *_ = orc.backend_raw(gates, node_bound, [], [])
    if st:
        return None
    level = np.zeros(G, dtype=np.int64)
    for g in order:
        l = 0
        for s in (1, 2):
            d = prod[gates[g, s]]
            if d >= 0:
                l = max(l, level[d] + 1)
        level[g] = l
    return level


def check_levels(ctx, orc, gates, node_bound):
    gates = np.asarray(gates, dtype=np.uint32).reshape(-1, 4)
    want = cpu_levels(orc, gates, node_bound)
    lo, off = ctx.topo_levels(gates, node_bound)
    G = gates.shape[0]
    assert off[0] == 0 and off[-1] == G and (np.diff(off.astype(np.int64)) > 0).all()
    assert np.array_equal(np.sort(lo), np.arange(G, dtype=np.uint32))
    got = np.empty(G, dtype=np.int64)
    for l in range(len(off) - 1):
        got[lo[off[l]:off[l + 1]]] = l
    assert np.array_equal(got, want)
    return len(off) - 1


@pytest.mark.parametrize("G,p_forward,window", [(1, 0, 0), (40, 1.0, 0), (3000, 0.3, 0), (20000, 1.0, 0), (20000, 0.0, 4), (50000, 0.5, 64)])
def test_levels_random(ctx, orc, c2a, G, p_forward, window):
    gates, nb = c2a.workloads.random_gates(G, 1 + G // 9, seed=G + window, p_forward=p_forward, p_dup_out=0.0, window=window)
    check_levels(ctx, orc, gates, nb)


def test_levels_empty(ctx):
    lo, off = ctx.topo_levels(np.zeros((0, 4), dtype=np.uint32), 4)
    assert lo.shape[0] == 0 and off.tolist() == [0]


def test_levels_deep_chain_small_frontier_mode(ctx, orc):
    """one chain of 6000 gates: 6000 levels of one gate each -> CTA 0 runs them without grid barriers"""
    G = 6000
    g = np.zeros((G, 4), dtype=np.uint32)
    g[:, 3] = 10 + np.arange(G)
    g[1:, 1] = 10 + np.arange(G - 1)
    g[0, 1] = 1
    g[:, 2] = 2
    assert check_levels(ctx, orc, g, 10 + G) == G


def test_levels_mixed_wide_and_narrow(ctx, orc, c2a):
    """frontier sizes alternate across the small/large threshold"""
    wl = c2a.workloads.mimc_chains(3000, rounds=6, variant="late")
    comp = c2a.Compiler(context=ctx)
    comp.emit_events(wl.events)
    nl = check_levels(ctx, orc, comp.gate_array(), comp.node_count + 1)
    assert nl == 6 * 6 + 1


def test_levels_high_fanout_rows_use_bulk_path(ctx, orc):
    """gate 0 feeds 9000 gates, gate 1 feeds 1500, gate 2 feeds 40 (medium), each consumer feeds one more gate"""
    rows = [(0, 9000), (1, 1500), (2, 40)]
    gl = [[7, 1, 2, 10], [7, 1, 2, 11], [7, 2, 2, 12]]
    node = 100
    for src, n in rows:
        for _ in range(n):
            gl.append([0, 10 + src, 10 + src if len(gl) % 3 == 0 else 1, node])   # some use the producer on both operands
            gl.append([0, node, 2, node + 1])
            node += 2
    check_levels(ctx, orc, np.array(gl, dtype=np.uint32), node + 1)


def test_levels_cycle_and_cap(ctx, c2a):
    A = 0
    g = np.array([[A, 1, 2, 3], [A, 12, 1, 10], [A, 10, 1, 11], [A, 11, 3, 12], [A, 3, 3, 13]], dtype=np.uint32)
    with pytest.raises(c2a.CircuitError) as e:
        ctx.topo_levels(g, 14)
    assert int(e.value.status) == 1 and e.value.message == "detected at i=1"   # smallest gate index left behind
    chain = np.zeros((50, 4), dtype=np.uint32)
    chain[:, 3] = 10 + np.arange(50)
    chain[1:, 1] = 10 + np.arange(49)
    chain[:, 2] = 1
    chain[0, 1] = 1
    with pytest.raises(c2a.C2AError):
        ctx.topo_levels(chain, 61, level_cap=10)
    lo, off = ctx.topo_levels(chain, 61, level_cap=50)
    assert len(off) == 51


def ref_exec(orc, op, a, b):
    st, r, _ = orc.execute_op(int(a), int(b), int(op))
    return (st == 0), r


def test_sweep_masks(ctx, orc, c2a):
    rng = np.random.RandomState(5)
    G, n_free = 4000, 60
    gates, nb = c2a.workloads.random_gates(G, n_free, seed=11, p_forward=0.4, p_dup_out=0.0, window=0)
    const_nodes = np.arange(1, 31, dtype=np.uint32)
    const_vals = rng.randint(0, 6, size=30).astype(np.uint32)
    outs = rng.choice(np.unique(gates[:, 3]), size=25, replace=False).astype(np.uint32)
    cm, cv, dm = ctx.sweep_masks(gates, nb, const_nodes, const_vals, outs)
    # CPU restatement
    st, _, order, *_ = orc.backend_raw(gates, nb, [], [])
    assert st == 0
    prod = np.full(nb, -1, dtype=np.int64)
    prod[gates[:, 3]] = np.arange(G)
    known = {int(n): int(v) for n, v in zip(const_nodes, const_vals)}
    want_c = np.zeros(G, dtype=np.uint8)
    want_v = np.zeros(G, dtype=np.uint32)
    for g in order:
        op, l, r, o = (int(x) for x in gates[g])
        ok = False
        if l in known and r in known:
            ok, val = ref_exec(orc, op, known[l], known[r])
        if ok:
            want_c[g], want_v[g] = 1, val
        if prod[o] == g:
            if ok:
                known[o] = val
            else:
                known.pop(o, None)
    assert np.array_equal(cm, want_c) and np.array_equal(cv, want_v)
    live = np.zeros(G, dtype=bool)
    outset = set(int(x) for x in outs)
    consumers = [[] for _ in range(G)]
    for g in range(G):
        for s in (1, 2):
            d = prod[gates[g, s]]
            if d >= 0:
                consumers[d].append(g)
    for g in order[::-1]:
        live[g] = (int(gates[g, 3]) in outset and prod[gates[g, 3]] == g) or any(live[c] for c in consumers[g])
    assert np.array_equal(dm.astype(bool), ~live)
    assert 0 < dm.sum() < G and 0 < cm.sum() < G
