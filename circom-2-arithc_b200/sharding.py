"""Multi-GPU build_circuit: shard the gate vector across ranks where the DAG splits into independent component
subtrees, build every shard with the single-GPU pipeline, and stitch the global numbering with ONE all-gather of
per-rank counts (+ the gathers that assemble the global arrays).  SURVEY.md §8e / DESIGN.md §6.

Why contiguous cuts are enough and exact.  The reference sorts with a DFS whose roots ascend (src/topological_sort.rs:
11-13) and numbers intermediate wires first-seen over the sorted gates (src/compiler.rs:427-443).  If no dependency
edge and no non-I/O node crosses a cut at gate index c, then
  * the DFS started from a root < c never reaches a gate >= c and vice versa  => global order = concat(per-shard order)
  * first-seen numbering of shard k starts exactly where shard k-1 stopped    => global wire = local wire + offset
  * inputs are numbered before (src/compiler.rs:392-395) and outputs after (:446-449) all intermediates, from lists that
    every rank shares, so they only shift by the other shards' intermediate counts.
A DAG without such cuts (one SHA-256 / Keccak instance, one long chain) does not shard: plan_shards returns None and the
caller runs replicas / a single GPU.

One process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).  The per-shard build and the wire rebase
are injected callables so that the host-side logic here is testable without a device; the defaults run on the GPU
through the C ABI and there is no CPU implementation of them in the product.
"""
from typing import Callable, List, Optional, Tuple

import numpy as np

NONE = 0xFFFFFFFF


def find_cuts(gates: np.ndarray, node_bound: int, input_nodes, output_nodes) -> np.ndarray:
    """Gate indices c (0 < c < G) such that gates[:c] and gates[c:] share no dependency edge and no non-I/O node."""
    g = np.ascontiguousarray(gates, dtype=np.uint32).reshape(-1, 4)
    G = g.shape[0]
    if G < 2:
        return np.zeros(0, dtype=np.int64)
    idx = np.arange(G, dtype=np.int64)
    # last producer per node (src/compiler.rs:401-406: HashMap::insert, the last gate wins)
    prod = np.full(node_bound, -1, dtype=np.int64)
    np.maximum.at(prod, g[:, 3], idx)
    diff = np.zeros(G + 2, dtype=np.int64)

    def forbid(lo, hi):  # cuts c with lo < c <= hi are invalid
        m = hi > lo
        np.add.at(diff, lo[m] + 1, 1)
        np.add.at(diff, hi[m] + 1, -1)

    for slot in (1, 2):  # dependency edges (src/compiler.rs:408-421)
        d = prod[g[:, slot]]
        m = d >= 0
        forbid(np.minimum(idx[m], d[m]), np.maximum(idx[m], d[m]))
    # non-I/O nodes must live in one shard (their first-seen wire id is local to it)
    first = np.full(node_bound, G, dtype=np.int64)
    last = np.full(node_bound, -1, dtype=np.int64)
    for slot in (1, 2, 3):
        np.minimum.at(first, g[:, slot], idx)
        np.maximum.at(last, g[:, slot], idx)
    io = np.zeros(node_bound, dtype=bool)
    io[np.asarray(input_nodes, dtype=np.int64)] = True
    io[np.asarray(output_nodes, dtype=np.int64)] = True
    used = (last >= 0) & ~io
    forbid(first[used], last[used])
    cover = np.cumsum(diff)[: G + 1]
    c = np.arange(1, G, dtype=np.int64)
    return c[cover[1:G] == 0]


def plan_shards(gates: np.ndarray, node_bound: int, input_nodes, output_nodes, world: int) -> Optional[List[Tuple[int, int]]]:
    """Contiguous gate ranges [(lo, hi)] * world balanced by gate count, or None when the DAG does not split `world` ways."""
    G = int(np.asarray(gates).reshape(-1, 4).shape[0])
    if world <= 1:
        return [(0, G)]
    cuts = find_cuts(gates, node_bound, input_nodes, output_nodes)
    if cuts.size < world - 1:
        return None
    bounds = [0]
    for k in range(1, world):
        target = k * G // world
        j = int(np.searchsorted(cuts, target))
        cand = [cuts[i] for i in (j - 1, j) if 0 <= i < cuts.size and cuts[i] > bounds[-1]]
        if not cand:
            later = cuts[cuts > bounds[-1]]
            if later.size == 0:
                return None
            cand = [later[0]]
        bounds.append(int(min(cand, key=lambda c: abs(int(c) - target))))
    bounds.append(G)
    if any(b <= a for a, b in zip(bounds, bounds[1:])):
        return None
    return list(zip(bounds[:-1], bounds[1:]))


def plan_shards_device(ctx, d_gates_ptr: int, G: int, node_bound: int, input_nodes, output_nodes, world: int) -> Optional[List[Tuple[int, int]]]:
    """plan_shards on the GPU (c2a_plan_shards_device): the gate vector is already resident at device address d_gates_ptr (e.g. a
    torch tensor's data_ptr()).  Same result as plan_shards() whenever a neighbouring cut exists for every target."""
    import ctypes as C
    from ._lib import lib
    ins = np.ascontiguousarray(input_nodes, dtype=np.uint32)
    outs = np.ascontiguousarray(output_nodes, dtype=np.uint32)
    bounds = (C.c_uint64 * (world + 1))()
    n = C.c_uint32(0)
    vp = C.c_void_p
    st = lib.c2a_plan_shards_device(ctx.handle, vp(d_gates_ptr), G, node_bound, ins.ctypes.data_as(vp), len(ins), outs.ctypes.data_as(vp), len(outs),
                                    world, bounds, C.byref(n))
    if st != 0:
        raise RuntimeError(f"c2a_plan_shards_device -> {st}: {ctx.last_error()}")
    if n.value != world:
        return None
    return [(int(bounds[k]), int(bounds[k + 1])) for k in range(world)]


def rebase_offsets(counts: np.ndarray, rank: int, shared_io: bool):
    """counts[r] = (n_in, n_mid, n_out, G) of rank r  ->  (off_in, off_mid, off_out, gate_base) for c2a_rebase_wires_device.
    shared_io=True : every rank was given the SAME global input/output lists (sharded build of one circuit);
    shared_io=False: every rank owns its own inputs/outputs (independent circuits side by side, bench.py weak scaling)."""
    counts = np.asarray(counts, dtype=np.int64).reshape(-1, 4)
    n_in, n_mid, n_out = (int(x) for x in counts[rank, :3])
    tot_mid = int(counts[:, 1].sum())
    gate_base = int(counts[:rank, 3].sum())
    if shared_io:
        return 0, int(counts[:rank, 1].sum()), tot_mid - n_mid, gate_base
    tot_in = int(counts[:, 0].sum())
    off_in = int(counts[:rank, 0].sum())
    off_mid = tot_in + int(counts[:rank, 1].sum()) - n_in
    off_out = tot_in + tot_mid + int(counts[:rank, 2].sum()) - n_in - n_mid
    return off_in, off_mid, off_out, gate_base


def _cuda_build_local(ctx):
    import ctypes as C
    import torch
    from ._lib import lib

    def build(gates_local: np.ndarray, node_bound: int, ins: np.ndarray, outs: np.ndarray):
        dev = torch.device("cuda", ctx.device)
        G = gates_local.shape[0]
        d_gates = torch.from_numpy(np.ascontiguousarray(gates_local).view(np.int32)).to(dev)
        d_order = torch.empty(G, dtype=torch.int32, device=dev)
        d_wire = torch.empty(node_bound, dtype=torch.int32, device=dev)
        d_new = torch.empty((G, 4), dtype=torch.int32, device=dev)
        wc, err, vp = C.c_uint32(0), C.c_uint64(0), C.c_void_p
        st = lib.c2a_build_circuit_device(ctx.handle, vp(d_gates.data_ptr()), G, node_bound, ins.ctypes.data_as(vp), len(ins),
                                          outs.ctypes.data_as(vp), len(outs), vp(d_order.data_ptr()), vp(d_wire.data_ptr()), vp(d_new.data_ptr()),
                                          C.byref(wc), C.byref(err))
        if st != 0:
            from .compiler import _raise
            _raise(st, f"detected at i={err.value}" if st == 1 else ctx.last_error())
        return d_order, d_wire, d_new, int(wc.value)

    def rebase(d_order, d_wire, d_new, n_in, n_mid, off_in, off_mid, off_out, gate_base):
        vp = C.c_void_p
        st = lib.c2a_rebase_wires_device(ctx.handle, vp(d_new.data_ptr()), vp(d_order.data_ptr()), d_order.shape[0], n_in, n_mid, off_in, off_mid, off_out, gate_base)
        if st == 0:
            st = lib.c2a_rebase_wire_map_device(ctx.handle, vp(d_wire.data_ptr()), d_wire.shape[0], n_in, n_mid, off_in, off_mid, off_out)
        if st != 0:
            raise RuntimeError(f"rebase -> {st}: {ctx.last_error()}")

    return build, rebase


def build_circuit_sharded(gates: np.ndarray, node_bound: int, input_nodes, output_nodes, *, ctx=None, group=None,
                          build_local: Optional[Callable] = None, rebase: Optional[Callable] = None):
    """Compiler::build_circuit's back end (src/compiler.rs:388-464) across the ranks of `group`.
    Every rank passes the SAME global gate vector / I/O lists and receives the global
    (order[G], wire_of_node[node_bound], new_gates[G,4], wire_count, plan).  Raises ValueError when the DAG does not
    split (the caller then runs single-GPU / replicas)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    g = np.ascontiguousarray(gates, dtype=np.uint32).reshape(-1, 4)
    ins = np.ascontiguousarray(input_nodes, dtype=np.uint32)
    outs = np.ascontiguousarray(output_nodes, dtype=np.uint32)
    plan = plan_shards(g, node_bound, ins, outs, world)
    if plan is None:
        raise ValueError("the gate DAG does not split into independent contiguous shards: replicas only")
    if build_local is None or rebase is None:
        build_local, rebase = _cuda_build_local(ctx)
    lo, hi = plan[rank]
    d_order, d_wire, d_new, wc = build_local(g[lo:hi], node_bound, ins, outs)
    dev = d_order.device
    n_mid = wc - len(ins) - len(outs)
    # the one exchange on the data path: per-rank (n_in, n_mid, n_out, G)
    mine = torch.tensor([len(ins), n_mid, len(outs), hi - lo], dtype=torch.int64, device=dev)
    allc = torch.empty(4 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allc, mine, group=group)
    counts = allc.view(world, 4).cpu().numpy()
    off_in, off_mid, off_out, gate_base = rebase_offsets(counts, rank, shared_io=True)
    assert gate_base == lo
    rebase(d_order, d_wire, d_new, len(ins), n_mid, off_in, off_mid, off_out, gate_base)
    # assemble the global arrays on every rank (padded all-gathers; the wire map is a max-reduce: a node is numbered by
    # at most one shard, I/O nodes identically by all, "no wire" is -1 as int32)
    Gmax = int(counts[:, 3].max())
    pad_o = torch.zeros(Gmax, dtype=torch.int32, device=dev)
    pad_g = torch.zeros((Gmax, 4), dtype=torch.int32, device=dev)
    pad_o[: hi - lo] = d_order
    pad_g[: hi - lo] = d_new
    all_o = torch.empty(world * Gmax, dtype=torch.int32, device=dev)
    all_g = torch.empty((world * Gmax, 4), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(all_o, pad_o, group=group)
    dist.all_gather_into_tensor(all_g, pad_g, group=group)
    dist.all_reduce(d_wire, op=dist.ReduceOp.MAX, group=group)
    order = np.concatenate([all_o[r * Gmax: r * Gmax + int(counts[r, 3])].cpu().numpy() for r in range(world)]).astype(np.uint32)
    new_gates = np.concatenate([all_g[r * Gmax: r * Gmax + int(counts[r, 3])].cpu().numpy() for r in range(world)]).astype(np.uint32)
    wire = d_wire.cpu().numpy().astype(np.uint32)
    wire_count = len(ins) + int(counts[:, 1].sum()) + len(outs)
    return order, wire, new_gates, wire_count, plan
