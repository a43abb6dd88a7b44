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
