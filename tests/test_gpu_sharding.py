"""2-GPU sharded build_circuit through the C ABI + NCCL against the oracle's single-process result.
Skipped on a 1-GPU box (run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
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
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ctx = c2a.DeviceContext(rank)
        ok = True
        for wl in (c2a.workloads.mimc_chains(400, rounds=91, variant="late"), c2a.workloads.keccak_shaped(instances=2, rounds=4)):
            comp = c2a.Compiler()
            comp.emit_events(wl.events)
            g, nb = comp.gate_array(), comp.node_count + 1
            ins = comp.signal_nodes(np.array(sorted(wl.inputs), dtype=np.uint32))
            outs = comp.signal_nodes(np.array(sorted(wl.outputs), dtype=np.uint32))
            before = ctx.kernel_launches()
            order, wire, ng, wc, plan = c2a.sharding.build_circuit_sharded(g, nb, ins, outs, ctx=ctx)
            st, _, o_order, o_wire, o_gates, o_wc = orc.backend_raw(g, nb, ins, outs)
            ok = ok and st == 0 and wc == o_wc and np.array_equal(order, o_order) and np.array_equal(wire, o_wire) and np.array_equal(ng, o_gates)
            ok = ok and ctx.kernel_launches() > before and len(plan) == world
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharded_build_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


@pytest.mark.parametrize("name", ["mimc_late", "mimc_inorder", "keccak2", "sha", "shared_node"])
def test_device_shard_planner_matches_the_host_planner(c2a, ctx, name):
    """c2a_plan_shards_device (first / last use per node, dependency spans, one running maximum on the GPU) against
    sharding.plan_shards (numpy on the host) - 1 GPU is enough: planning is per rank"""
    import torch
    if name == "shared_node":   # two otherwise independent gates share the producer-less, non-I/O node 9
        cases = [(np.array([[0, 1, 9, 5], [0, 2, 9, 6]], dtype=np.uint32), 10, [1, 2], [5, 6]),
                 (np.array([[0, 1, 9, 5], [0, 2, 9, 6]], dtype=np.uint32), 10, [1, 2, 9], [5, 6])]
    else:
        wl = {"mimc_late": lambda: c2a.workloads.mimc_chains(61, rounds=17, variant="late"),
              "mimc_inorder": lambda: c2a.workloads.mimc_chains(40, rounds=9, variant="inorder"),
              "keccak2": lambda: c2a.workloads.keccak_shaped(instances=2, rounds=2),
              "sha": lambda: c2a.workloads.sha256_shaped(rounds=3)}[name]()
        comp = c2a.Compiler(context=ctx)
        comp.emit_events(wl.events)
        cases = [(comp.gate_array(), comp.node_count + 1, comp.signal_nodes(np.array(sorted(wl.inputs), dtype=np.uint32)),
                  comp.signal_nodes(np.array(sorted(wl.outputs), dtype=np.uint32)))]
    for g, nb, ins, outs in cases:
        d = torch.from_numpy(np.ascontiguousarray(g).view(np.int32)).to("cuda:0")
        for world in (1, 2, 3, 4, 8):
            want = c2a.sharding.plan_shards(g, nb, ins, outs, world)
            got = c2a.sharding.plan_shards_device(ctx, d.data_ptr(), g.shape[0], nb, ins, outs, world)
            assert got == want, (name, world, got, want)
