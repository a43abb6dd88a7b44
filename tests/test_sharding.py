"""Multi-GPU host logic on CPU: shard planning, offset arithmetic, and the whole sharded build over a
world_size-2 gloo group where the per-shard build is the ORACLE's back end (test infrastructure standing in for the
CUDA pipeline, which cannot run here) — the stitched result must equal the oracle's single-process build_circuit."""
import os
import socket

import numpy as np
import pytest


def _circuit(c2a, wl):
    comp = c2a.Compiler()
    comp.emit_events(wl.events)
    g = comp.gate_array()
    ins = comp.signal_nodes(np.array(sorted(wl.inputs), dtype=np.uint32))
    outs = comp.signal_nodes(np.array(sorted(wl.outputs), dtype=np.uint32))
    return g, comp.node_count + 1, ins, outs


def test_cuts_are_chain_boundaries(c2a):
    wl = c2a.workloads.mimc_chains(6, rounds=5, variant="late")
    g, nb, ins, outs = _circuit(c2a, wl)
    cuts = c2a.sharding.find_cuts(g, nb, ins, outs)
    per = g.shape[0] // 6
    assert cuts.tolist() == [per * k for k in range(1, 6)]
    assert c2a.sharding.plan_shards(g, nb, ins, outs, 2) == [(0, 3 * per), (3 * per, 6 * per)]
    assert c2a.sharding.plan_shards(g, nb, ins, outs, 4) == [(0, 2 * per), (2 * per, 3 * per), (3 * per, 5 * per), (5 * per, 6 * per)] or \
        len(c2a.sharding.plan_shards(g, nb, ins, outs, 4)) == 4
    assert c2a.sharding.plan_shards(g, nb, ins, outs, 8) is None  # only 6 independent subtrees


def test_single_component_does_not_shard(c2a):
    wl = c2a.workloads.sha256_shaped(rounds=4)
    g, nb, ins, outs = _circuit(c2a, wl)
    assert c2a.sharding.find_cuts(g, nb, ins, outs).size == 0
    assert c2a.sharding.plan_shards(g, nb, ins, outs, 2) is None
    wl = c2a.workloads.keccak_shaped(instances=2, rounds=1)
    g, nb, ins, outs = _circuit(c2a, wl)
    assert c2a.sharding.find_cuts(g, nb, ins, outs).tolist() == [g.shape[0] // 2]   # two sponges: exactly one cut


def test_shared_intermediate_node_blocks_a_cut(c2a):
    # two otherwise independent gates share the producer-less, non-I/O node 9 (e.g. a constant of the enclosing context)
    g = np.array([[0, 1, 9, 5], [0, 2, 9, 6]], dtype=np.uint32)
    assert c2a.sharding.find_cuts(g, 10, [1, 2], [5, 6]).size == 0
    assert c2a.sharding.find_cuts(g, 10, [1, 2, 9], [5, 6]).tolist() == [1]   # as an input it is numbered up front: fine


def test_rebase_offsets(c2a):
    counts = np.array([[3, 10, 2, 100], [3, 7, 2, 50], [3, 5, 2, 70]])
    assert c2a.sharding.rebase_offsets(counts, 1, shared_io=True) == (0, 10, 22 - 7, 100)
    assert c2a.sharding.rebase_offsets(counts, 0, shared_io=True) == (0, 0, 12, 0)
    # independent circuits side by side: inputs of all ranks first, then all intermediates, then all outputs
    assert c2a.sharding.rebase_offsets(counts, 1, shared_io=False) == (3, 9 + 10 - 3, 9 + 22 + 2 - 3 - 7, 100)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, variant, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch
    import torch.distributed as dist
    from c2a_loader import c2a
    import oracle_lib as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wl = c2a.workloads.mimc_chains(9, rounds=7, variant=variant) if variant != "keccak" else c2a.workloads.keccak_shaped(instances=2, rounds=2)
        g, nb, ins, outs = _circuit(c2a, wl)

        def build_local(gl, node_bound, i, o):   # the oracle's back end stands in for c2a_build_circuit_device
            st, _, order, wire, ng, wc = orc.backend_raw(gl, node_bound, i, o)
            assert st == 0
            t = lambda a: torch.from_numpy(a.view(np.int32).copy())
            return t(order), t(wire), t(ng), wc

        def rebase(d_order, d_wire, d_new, n_in, n_mid, off_in, off_mid, off_out, gate_base):   # restates k_rebase / k_rebase_map
            def fix(w):
                w = w.to(torch.int64) & 0xFFFFFFFF
                r = torch.where(w < n_in, w + off_in, torch.where(w < n_in + n_mid, w + off_mid, w + off_out))
                return torch.where(w == 0xFFFFFFFF, w, r)
            d_new[:, 1:] = fix(d_new[:, 1:]).to(torch.int32)
            d_wire.copy_(fix(d_wire).to(torch.int32))
            d_order += gate_base

        order, wire, ng, wc, plan = c2a.sharding.build_circuit_sharded(g, nb, ins, outs, build_local=build_local, rebase=rebase)
        st, _, o_order, o_wire, o_gates, o_wc = orc.backend_raw(g, nb, ins, outs)
        ok = st == 0 and wc == o_wc and np.array_equal(order, o_order) and np.array_equal(wire, o_wire) and np.array_equal(ng, o_gates) and len(plan) == world
        q.put((rank, bool(ok), plan))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("variant", ["late", "inorder", "keccak"])
def test_sharded_build_world2_gloo(variant):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, variant, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == res[1][2]
