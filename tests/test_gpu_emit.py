"""GPU parity of the DEVICE emitter (csrc/c2a_emit.cuh, through the C ABI): replaying a whole event stream on the GPU
must give the reference's node ids and node-id gate vector bit for bit (src/compiler.rs:139-278), and the resident
result must build the same circuit as the oracle's build_circuit (src/compiler.rs:321-494).
Small streams are checked against the CPU oracle (faithful linear scans); large ones against the product's host
union-find emitter, which tests/test_host_emitter.py pins to the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EV_S, EV_SC, EV_G, EV_C = 0, 1, 2, 3
DEVICE, HOST = 1, 2


class _DSU:
    def __init__(self):
        self.p, self.out, self.const = {}, {}, {}

    def add(self, s, const):
        self.p[s] = s
        self.out[s] = False
        self.const[s] = const

    def find(self, x):
        while self.p[x] != x:
            self.p[x] = self.p[self.p[x]]
            x = self.p[x]
        return x

    def can_merge(self, a, b):
        ra, rb = self.find(a), self.find(b)
        return ra == rb or not ((self.out[ra] and self.out[rb]) or (self.const[ra] and self.const[rb]))

    def merge(self, a, b):
        ra, rb = self.find(a), self.find(b)
        if ra != rb:
            self.p[rb] = ra
            self.out[ra] |= self.out[rb]
            self.const[ra] |= self.const[rb]


def valid_stream(rng, n_events, id_order="sequential", p_redundant=0.15, p_const=0.1, shape="random"):
    """A stream the reference accepts: every reference is declared earlier, every gate writes a fresh temporary
    (src/process.rs:470-475), no out+out / const+const merge.  Contains redundant connections (both signals already
    in one node), self connections, and - by `shape` - long connection paths / stars that need several Boruvka rounds."""
    n_ids = n_events + 8
    ids = np.arange(n_ids)
    if id_order == "shuffled":
        ids = rng.permutation(n_ids)
    elif id_order == "gappy":
        ids = np.sort(rng.choice(np.arange(3 * n_ids), size=n_ids, replace=False))
    nxt = [0]

    def fresh():
        i = int(ids[nxt[0]])
        nxt[0] += 1
        return i

    d = _DSU()
    ev, declared, plain = [], [], []
    for _ in range(3):
        s = fresh()
        d.add(s, False)
        ev.append((EV_S, s, 0, 0))
        declared.append(s)
        plain.append(s)
    while len(ev) < n_events:
        x = rng.rand()
        if x < 0.30:
            s = fresh()
            const = rng.rand() < p_const
            d.add(s, const)
            ev.append((EV_SC, s, int(rng.randint(0, 1000)), 0) if const else (EV_S, s, 0, 0))
            declared.append(s)
            if not const:
                plain.append(s)
        elif x < 0.60:
            o = fresh()
            d.add(o, False)
            ev.append((EV_S, o, 0, 0))
            a, b = (int(declared[rng.randint(len(declared))]) for _ in range(2))
            ev.append((EV_G | (int(rng.randint(0, 20)) << 8), a, b, o))
            d.out[d.find(o)] = True
            declared.append(o)
        else:
            if shape == "path" and len(plain) > 1:      # extend one long class: plain[k] -- plain[k+1] in random order
                k = int(rng.randint(len(plain) - 1))
                a, b = plain[k], plain[k + 1]
            elif shape == "star":
                a, b = plain[0], int(declared[rng.randint(len(declared))])
            else:
                a, b = (int(declared[rng.randint(len(declared))]) for _ in range(2))
            if rng.rand() < p_redundant and len(ev) > 10:  # re-connect two signals that already share a node
                r = d.find(a)
                same = [s for s in declared[-50:] if d.find(s) == r]
                b = int(same[rng.randint(len(same))]) if same else a
            if d.can_merge(a, b):
                d.merge(a, b)
                ev.append((EV_C, a, b, 0))
    return np.asarray(ev, dtype=np.uint32).reshape(-1, 4)


def oracle_emit(orc, ev):
    oc = orc.OracleCompiler()
    oc.emit_events(ev)
    return oc


def check_against(ctx, ref, ev, expect_path=DEVICE):
    """ref: an emitter with gate_array() / node_count / signal_node() (the oracle or the product's host emitter)."""
    info = ctx.emit_events(ev)
    assert info["path"] == expect_path, info
    gates, nos = ctx.emitted_fetch()
    rg = ref.gate_array()
    assert info["n_gates"] == rg.shape[0]
    assert info["node_count"] == ref.node_count
    assert np.array_equal(gates, rg), f"gate vector differs first at {np.argmax((gates != rg).any(axis=1))}"
    kinds = ev[:, 0] & 0xFF
    sids = ev[kinds <= 1, 1]
    if hasattr(ref, "signal_nodes"):
        want = ref.signal_nodes(sids)
    else:
        want = np.array([ref.signal_node(int(s)) for s in sids], dtype=np.uint32)
    assert np.array_equal(nos[sids], want)
    undeclared = np.setdiff1d(np.arange(info["signal_bound"]), sids)
    assert (nos[undeclared] == 0).all()
    # the packed form of the same stream (kinds byte + payload words) must replay to the identical result
    from c2a_loader import c2a
    kinds_b, words, flags = c2a.pack_events(ev)
    assert bool(flags & 1) == bool(np.array_equal(sids, np.arange(len(sids))))  # C2A_PACKED_DENSE_IDS
    info_p = ctx.emit_packed(kinds_b, words, flags)
    gates_p, nos_p = ctx.emitted_fetch()
    assert {k: v for k, v in info_p.items() if k != "decline_flags"} == {k: v for k, v in info.items() if k != "decline_flags"}
    assert np.array_equal(gates_p, gates) and np.array_equal(nos_p, nos)
    if flags & 1:   # dense ids: the 4 B/event form (operands derived from the signal declared last carry no word) replays identically
        kinds_i, words_i, flags_i = c2a.pack_events(ev, implicit=True)
        assert flags_i == 3 and len(words_i) <= len(words)
        info_i = ctx.emit_packed(kinds_i, words_i, flags_i)
        gates_i, nos_i = ctx.emitted_fetch()
        assert {k: v for k, v in info_i.items() if k != "decline_flags"} == {k: v for k, v in info.items() if k != "decline_flags"}
        assert np.array_equal(gates_i, gates) and np.array_equal(nos_i, nos)
    return info, gates, nos


@pytest.mark.parametrize("seed", range(24))
def test_valid_random_streams_match_oracle(ctx, orc, seed):
    rng = np.random.RandomState(seed)
    n = int(rng.choice([12, 40, 300, 2000, 5000]))
    ev = valid_stream(rng, n, id_order=["sequential", "shuffled", "gappy"][seed % 3], shape=["random", "path", "star"][(seed // 3) % 3],
                      p_redundant=[0.0, 0.15, 0.5][seed % 3])
    info, _, _ = check_against(ctx, oracle_emit(orc, ev), ev)
    assert info["n_effective"] <= info["n_connections"]


def test_long_connection_path_needs_many_rounds(ctx, orc):
    """one class built from a 4096-signal path whose connections arrive in random order: Boruvka needs ~log2 rounds,
    and the node ids depend on the exact arrival order"""
    rng = np.random.RandomState(7)
    n = 4096
    ev = [(EV_S, i, 0, 0) for i in range(n)]
    for k in rng.permutation(n - 1):
        ev.append((EV_C, int(k), int(k) + 1, 0))
    ev += [(EV_C, 5, 900, 0), (EV_C, 17, 17, 0)]  # redundant
    ev = np.asarray(ev, dtype=np.uint32)
    info, _, nos = check_against(ctx, oracle_emit(orc, ev), ev)
    assert info["rounds"] >= 5 and info["n_effective"] == n - 1
    assert len(set(nos[:n].tolist())) == 1


def test_many_rounds_with_gates_then_build(ctx, orc):
    """gates over a class that needs more Boruvka rounds than the speculative ones: the node ids are computed twice, so the
    producer map the gate kernel fills for the build (compiler.rs:401-406) must be the second run's; build twice (it is kept)"""
    rng = np.random.RandomState(11)
    n, ng = 2048, 600
    ev = [(EV_S, i, 0, 0) for i in range(n)]
    pool = list(range(n))
    for j in range(ng):
        o = n + j
        ev.append((EV_S, o, 0, 0))
        a, b = (int(pool[rng.randint(len(pool))]) for _ in range(2))
        ev.append((EV_G | (int(rng.randint(0, 20)) << 8), a, b, o))
        pool.append(o)
    for k in rng.permutation(n - 1):
        ev.append((EV_C, int(k), int(k) + 1, 0))
    ev = np.asarray(ev, dtype=np.uint32)
    info, gates, nos = check_against(ctx, oracle_emit(orc, ev), ev)
    assert info["rounds"] > 2
    ins, outs = np.array([0], dtype=np.uint32), np.array([n + ng - 1], dtype=np.uint32)
    nb = info["node_count"] + 1
    st, _, o_order, o_wire, o_gates, o_wc = orc.backend_raw(gates, nb, nos[ins], nos[outs])
    assert st == 0
    for _ in range(2):
        order, wire, g2, wc = ctx.emitted_build_circuit(ins, outs)
        assert wc == o_wc and np.array_equal(order, o_order) and np.array_equal(wire, o_wire) and np.array_equal(g2, o_gates)


def test_tag_wraparound_many_rounds(ctx, c2a):
    """a 2^17-signal path whose connections arrive ordered by the number of trailing zeros of their position: every
    Boruvka round only pairs up neighbouring classes, so it takes 17 rounds (more than 7: the best[] tags wrap)"""
    n = 1 << 17
    ev = np.zeros((2 * n - 1, 4), dtype=np.uint32)
    ev[:n, 0], ev[:n, 1] = EV_S, np.arange(n)
    k = np.arange(n - 1, dtype=np.int64)
    tz = np.array([((int(x) + 1) & -(int(x) + 1)).bit_length() - 1 for x in k])
    k = k[np.lexsort((k, tz))].astype(np.uint32)
    ev[n:, 0], ev[n:, 1], ev[n:, 2] = EV_C, k, k + 1
    comp = c2a.Compiler()
    comp.emit_events(ev)
    info, _, _ = check_against(ctx, comp, ev)
    assert info["rounds"] > 7


def test_empty_and_tiny(ctx, orc):
    ev = np.zeros((0, 4), dtype=np.uint32)
    info = ctx.emit_events(ev)
    assert info["n_gates"] == 0 and info["node_count"] == 0
    order, wire, ng, wc = ctx.emitted_build_circuit([], [])
    assert wc == 0 and len(order) == 0
    ev = np.asarray([(EV_S, 0, 0, 0), (EV_S, 1, 0, 0), (EV_S, 2, 0, 0), (EV_S, 3, 0, 0), (EV_G | (0 << 8), 0, 1, 3), (EV_C, 3, 2, 0)], dtype=np.uint32)
    info, gates, nos = check_against(ctx, oracle_emit(orc, ev), ev)
    assert gates.tolist() == [[0, 1, 2, 5]] and info["node_count"] == 5  # SURVEY.md §4 golden `sum`
    order, wire, ng, wc = ctx.emitted_build_circuit([0, 1], [2])
    assert wc == 3 and ng.tolist() == [[0, 0, 1, 2]]


@pytest.mark.parametrize("seed", range(30))
def test_streams_the_device_declines_replay_exactly(ctx, orc, c2a, seed):
    """duplicates, unknown references (node 0 semantics), merge errors: same status and event index as the oracle,
    same circuit when the reference accepts the stream"""
    from test_host_emitter import random_stream
    rng = np.random.RandomState(1000 + seed)
    raw = random_stream(rng, n_events=int(rng.randint(20, 300)), n_ids=int(rng.randint(5, 80)), p_unknown=0.05 if seed % 2 else 0.0)
    ev = []
    for e in raw:
        if e[0] == "S":
            ev.append((EV_SC, e[1], e[2], 0) if e[2] is not None else (EV_S, e[1], 0, 0))
        elif e[0] == "G":
            ev.append((EV_G | (e[1] << 8), e[2], e[3], e[4]))
        else:
            ev.append((EV_C, e[1], e[2], 0))
    ev = np.asarray(ev, dtype=np.uint32)
    oc = orc.OracleCompiler()
    try:
        oc.emit_events(ev)
        err = None
    except orc.OracleError as e:
        err = e
    if err is None:
        info = ctx.emit_events(ev)
        gates, nos = ctx.emitted_fetch()
        assert np.array_equal(gates, oc.gate_array()) and info["node_count"] == oc.node_count
    else:
        with pytest.raises((c2a.CircuitError, c2a.C2AError)) as ex:
            ctx.emit_events(ev)
        assert int(ex.value.status) == err.status
        assert f"event {ex.value.err_event}" == err.message
        with pytest.raises((c2a.CircuitError, c2a.C2AError)) as ex:  # same through the packed stream
            ctx.emit_packed(*c2a.pack_events(ev))
        assert int(ex.value.status) == err.status
        assert f"event {ex.value.err_event}" == err.message


@pytest.mark.parametrize("seed", range(12))
def test_dense_packed_streams_with_forward_references(ctx, orc, c2a, seed):
    """DENSE packed streams are validated inside the scatter (id < signals declared so far).  References to signals that are
    declared LATER (node-0 semantics for operands, compiler.rs:183; panic for the out signal, :201) must still come out
    exactly like the reference: the device declines and the host emitter replays the unpacked stream."""
    rng = np.random.RandomState(4000 + seed)
    n_sig_total = int(rng.randint(6, 60))
    ev, declared = [], 0
    while declared < n_sig_total:
        x = rng.rand()
        if x < 0.45 or declared < 3:
            ev.append((EV_S, declared, 0, 0))
            declared += 1
        elif x < 0.8:
            hi = declared + (3 if rng.rand() < 0.25 else 0)          # sometimes reach past what is declared
            a, b = (int(rng.randint(0, min(hi, n_sig_total))) for _ in range(2))
            ev.append((EV_S, declared, 0, 0))                           # fresh out signal
            o = declared if rng.rand() < 0.9 else min(declared + 2, n_sig_total - 1)
            declared += 1
            ev.append((EV_G | (int(rng.randint(0, 20)) << 8), a, b, o))
        else:
            hi = declared + (2 if rng.rand() < 0.2 else 0)
            a, b = (int(rng.randint(0, min(hi, n_sig_total))) for _ in range(2))
            ev.append((EV_C, a, b, 0))
    ev = np.asarray(ev, dtype=np.uint32)
    kinds_b, words, flags = c2a.pack_events(ev)
    assert flags == 1
    oc = orc.OracleCompiler()
    try:
        oc.emit_events(ev)
        err = None
    except orc.OracleError as e:
        err = e
    if err is None:
        info = ctx.emit_packed(kinds_b, words, flags)
        gates, _ = ctx.emitted_fetch()
        assert np.array_equal(gates, oc.gate_array()) and info["node_count"] == oc.node_count
    else:
        for imp in (False, True):
            with pytest.raises((c2a.CircuitError, c2a.C2AError)) as ex:
                ctx.emit_packed(*c2a.pack_events(ev, implicit=imp))
            assert int(ex.value.status) == err.status
            assert f"event {ex.value.err_event}" == err.message


def test_implicit_operand_stream_at_scale(ctx, c2a):
    """1 M gates: the 4 B/event packed form (TMA-staged scatter with the third rank) against the 6 B/event form"""
    wl = c2a.workloads.mimc_chains(1832, rounds=91, variant="late")
    ev = np.ascontiguousarray(wl.events)
    k6, w6, f6 = c2a.pack_events(ev)
    k4, w4, f4 = c2a.pack_events(ev, implicit=True)
    assert f4 == 3 and (k4.nbytes + w4.nbytes) / len(ev) < 4.1 < (k6.nbytes + w6.nbytes) / len(ev)
    i6 = ctx.emit_packed(k6, w6, f6)
    g6, n6 = ctx.emitted_fetch()
    i4 = ctx.emit_packed(k4, w4, f4)
    g4, n4 = ctx.emitted_fetch()
    assert i4["path"] == DEVICE and {k: v for k, v in i4.items()} == {k: v for k, v in i6.items()}
    assert np.array_equal(g4, g6) and np.array_equal(n4, n6)


def test_merge_errors(ctx, c2a):
    S = lambda i: (EV_S, i, 0, 0)
    K = lambda i, v: (EV_SC, i, v, 0)
    # const + const (compiler.rs:243-245)
    ev = np.asarray([K(0, 1), K(1, 2), S(2), (EV_C, 0, 2, 0), (EV_C, 2, 1, 0)], dtype=np.uint32)
    with pytest.raises(c2a.CircuitError) as ex:
        ctx.emit_events(ev)
    assert ex.value.status == c2a.Status.CANNOT_MERGE_CONSTANT_NODES and ex.value.err_event == 4
    # out + out (compiler.rs:239-241)
    ev = np.asarray([S(0), S(1), S(2), S(3), (EV_G, 0, 1, 2), (EV_G, 0, 1, 3), (EV_C, 2, 3, 0)], dtype=np.uint32)
    with pytest.raises(c2a.CircuitError) as ex:
        ctx.emit_events(ev)
    assert ex.value.status == c2a.Status.CANNOT_MERGE_OUTPUT_NODES and ex.value.err_event == 6
    # duplicate declaration (compiler.rs:146-148)
    ev = np.asarray([S(0), S(1), S(0)], dtype=np.uint32)
    with pytest.raises(c2a.CircuitError) as ex:
        ctx.emit_events(ev)
    assert ex.value.status == c2a.Status.SIGNAL_ALREADY_DECLARED and ex.value.err_event == 2
    # two gate outputs in one class is legal when the second gate comes AFTER the merge: the conservative screen
    # declines and the host emitter accepts
    ev = np.asarray([S(0), S(1), S(2), S(3), (EV_G, 0, 1, 2), (EV_C, 2, 3, 0), (EV_G, 0, 1, 3)], dtype=np.uint32)
    info = ctx.emit_events(ev)
    assert info["path"] == HOST and info["decline_flags"] != 0
    gates, _ = ctx.emitted_fetch()
    assert gates.tolist() == [[0, 1, 2, 5], [0, 1, 2, 5]]


@pytest.mark.parametrize("dense", [True, False])
def test_packed_resident_with_unaligned_device_pointers(ctx, c2a, dense):
    """c2a_emit_packed_resident takes the caller's device pointers as they are: when they are not 16-byte aligned the TMA bulk
    staging is skipped (plain loads) and the result must not change"""
    import ctypes as C
    import torch
    from circom_2_arithc_b200._lib import lib, EmitInfo, PackedEvents
    wl = c2a.workloads.mimc_chains(40, rounds=9, variant="late")
    ev = np.ascontiguousarray(wl.events).copy()
    if not dense:  # explicit ids: renumber the signals with gaps
        ev_kind = ev[:, 0] & 0xFF
        remap = np.arange(int(ev[ev_kind <= 1, 1].max()) + 1, dtype=np.uint32) * 2 + 3
        ev[ev_kind <= 1, 1] = remap[ev[ev_kind <= 1, 1]]
        for col in (1, 2, 3):
            ev[ev_kind == 2, col] = remap[ev[ev_kind == 2, col]]
        for col in (1, 2):
            ev[ev_kind == 3, col] = remap[ev[ev_kind == 3, col]]
    kinds_b, words, flags = c2a.pack_events(ev)
    assert bool(flags & 1) == dense
    want = ctx.emit_packed(kinds_b, words, flags)
    g_want, nos_want = ctx.emitted_fetch()
    dev = torch.device("cuda", 0)
    dk = torch.zeros(len(kinds_b) + 64, dtype=torch.uint8, device=dev)
    dw = torch.zeros(len(words) + 64, dtype=torch.int32, device=dev)
    for ko, wo in ((0, 0), (1, 1), (3, 2), (16, 5)):
        dk[ko:ko + len(kinds_b)] = torch.from_numpy(kinds_b).to(dev)
        dw[wo:wo + len(words)] = torch.from_numpy(words.view(np.int32)).to(dev)
        pk = PackedEvents(dk.data_ptr() + ko, dw.data_ptr() + 4 * wo, len(kinds_b), len(words), flags, 0)
        info, bad = EmitInfo(), C.c_uint64(0)
        assert lib.c2a_emit_packed_resident(ctx.handle, C.byref(pk), C.byref(info), C.byref(bad)) == 0, ctx.last_error()
        assert info.path == DEVICE and info.node_count == want["node_count"] and info.n_gates == want["n_gates"]
        ctx._emit_info = {"n_gates": int(info.n_gates), "signal_bound": int(info.signal_bound)}
        g, nos = ctx.emitted_fetch()
        assert np.array_equal(g, g_want) and np.array_equal(nos, nos_want), (ko, wo)


def test_signal_wires_need_a_built_circuit(ctx, c2a):
    ev = np.asarray([(EV_S, 0, 0, 0), (EV_S, 1, 0, 0), (EV_S, 2, 0, 0), (EV_G, 0, 1, 2)], dtype=np.uint32)
    ctx.emit_events(ev)
    with pytest.raises(c2a.C2AError):
        ctx.emitted_signal_wires([0, 1])           # emitted, not built yet
    ctx.emitted_build_circuit([0, 1], [2])
    assert ctx.emitted_signal_wires([2, 0, 1, 7]).tolist() == [2, 0, 1, 0xFFFFFFFF]
    ctx.emit_events(ev)                            # a new emit drops the wire map
    with pytest.raises(c2a.C2AError):
        ctx.emitted_signal_wires([0])


def test_packed_stream_rejects_inconsistent_word_count(ctx, c2a):
    ev = np.asarray([(EV_S, 0, 0, 0), (EV_S, 1, 0, 0), (EV_S, 2, 0, 0), (EV_G, 0, 1, 2), (EV_C, 2, 1, 0)], dtype=np.uint32)
    kinds_b, words, flags = c2a.pack_events(ev)
    assert flags == 1 and words.tolist() == [0, 1, 2, 2, 1] and kinds_b.tolist() == [0, 0, 0, 2, 3]
    with pytest.raises(c2a.C2AError):
        ctx.emit_packed(kinds_b, words[:-1], flags)
    with pytest.raises(c2a.C2AError):
        ctx.emit_packed(kinds_b, np.concatenate([words, words[:1]]), flags)
    with pytest.raises(c2a.C2AError):
        ctx.emit_packed(kinds_b, words, 0)  # without DENSE_IDS the signals would need an id word each
    bad = kinds_b.copy()
    bad[3] = 2 | (25 << 2)                  # gate type out of range (src/a_gate_type.rs has 20)
    with pytest.raises((c2a.C2AError, c2a.CircuitError)):
        ctx.emit_packed(bad, words, flags)
    info = ctx.emit_packed(kinds_b, words, flags)
    assert info["path"] == DEVICE and info["n_gates"] == 1 and info["node_count"] == 4


def test_sparse_ids_go_through_the_host_emitter(ctx, c2a):
    ids = [0, 5, 4_000_000_000, 123_456_789, 7]
    ev = [(EV_S, i, 0, 0) for i in ids] + [(EV_G, ids[2], ids[3], ids[4]), (EV_C, ids[2], ids[0], 0)]
    ev = np.asarray(ev, dtype=np.uint32)
    info = ctx.emit_events(ev)
    assert info["path"] == HOST
    gates, _ = ctx.emitted_fetch(want_nodes=False)
    assert gates.tolist() == [[0, 6, 4, 5]]
    order, wire, ng, wc = ctx.emitted_build_circuit([ids[3]], [ids[4]])
    assert wc == 3 and ng.tolist() == [[0, 1, 0, 2]]


def _workload_case(ctx, orc, c2a, wl, vs_oracle_emit):
    ev = np.ascontiguousarray(wl.events)
    if vs_oracle_emit:
        ref = oracle_emit(orc, ev)
    else:
        ref = c2a.Compiler()
        ref.emit_events(ev)
    info, gates, nos = check_against(ctx, ref, ev)
    ins = np.array(sorted(wl.inputs), dtype=np.uint32)
    outs = np.array(sorted(wl.outputs), dtype=np.uint32)
    order, wire, ng, wc = ctx.emitted_build_circuit(ins, outs)
    nb = info["node_count"] + 1
    st, _, o_order, o_wire, o_gates, o_wc = orc.backend_raw(gates, nb, nos[ins], nos[outs])
    assert st == 0 and wc == o_wc
    assert np.array_equal(order, o_order) and np.array_equal(wire, o_wire) and np.array_equal(ng, o_gates)
    # named-wire lookups without the whole wire map (compiler.rs:323-383, 466-493): gates only + selected signals
    kinds = ev[:, 0] & 0xFF
    consts = ev[kinds == 1, 1][:1000]
    o2, w2, g2, wc2 = ctx.emitted_build_circuit(ins, outs, want_order=False, want_wires=False)
    assert o2 is None and w2 is None and wc2 == wc and np.array_equal(g2, ng)
    probe = np.concatenate([ins, outs, consts, np.array([info["signal_bound"] + 5], dtype=np.uint32)]).astype(np.uint32)
    got = ctx.emitted_signal_wires(probe)
    assert np.array_equal(got[:-1], o_wire[nos[probe[:-1]]]) and got[-1] == 0xFFFFFFFF
    return info


@pytest.mark.parametrize("variant", ["inorder", "late"])
def test_mimc_small_vs_oracle_emit(ctx, orc, c2a, variant):
    _workload_case(ctx, orc, c2a, c2a.workloads.mimc_chains(7, rounds=11, variant=variant), True)


def test_poseidon_vs_oracle_emit(ctx, orc, c2a):
    info = _workload_case(ctx, orc, c2a, c2a.workloads.poseidon_shaped(), True)
    assert info["n_gates"] == 1413


def test_sha256_keccak_vs_host_emitter(ctx, orc, c2a):
    _workload_case(ctx, orc, c2a, c2a.workloads.sha256_shaped(), False)
    _workload_case(ctx, orc, c2a, c2a.workloads.keccak_shaped(instances=2), False)


@pytest.mark.parametrize("variant", ["inorder", "late"])
def test_one_million_gates_device_emit(ctx, orc, c2a, variant):
    info = _workload_case(ctx, orc, c2a, c2a.workloads.mimc_chains(1832, rounds=91, variant=variant), False)
    assert info["n_gates"] >= 1_000_000 and info["path"] == DEVICE


def test_resident_circuit_survives_other_sizes(ctx, orc, c2a):
    """emit a small circuit, then a larger one (slab regrowth), then build: the resident result must be the last one"""
    a = c2a.workloads.mimc_chains(3, rounds=5, variant="late")
    b = c2a.workloads.mimc_chains(400, rounds=91, variant="late")
    ctx.emit_events(np.ascontiguousarray(a.events))
    _workload_case(ctx, orc, c2a, b, False)
    _workload_case(ctx, orc, c2a, a, True)


@pytest.mark.parametrize("name", ["sha256_full", "keccak_r8"])
def test_baseline_sized_streams_against_the_oracle_emit(ctx, orc, c2a, name):
    """BASELINE config 3 at FULL size (SHA-256-shaped, 115 920 gates, 233 K events) and a third of config 4 (Keccak-shaped, 8 rounds,
    64 K gates): the device emitter, the 4 B/event stream, the fused / multi-kernel compile against the ORACLE's faithful linear-scan
    emit (src/compiler.rs:139-278: about two minutes of host time for the SHA stream) and its back end - not against this
    repo's own host emitter."""
    wl = c2a.workloads.sha256_shaped() if name == "sha256_full" else c2a.workloads.keccak_shaped(1, rounds=8)
    ev = np.ascontiguousarray(wl.events)
    oc = orc.OracleCompiler()
    oc.emit_events(ev)
    ins, outs = np.array(sorted(wl.inputs), dtype=np.uint32), np.array(sorted(wl.outputs), dtype=np.uint32)
    nodes = lambda s: np.array([oc.signal_node(int(x)) for x in s], dtype=np.uint32)
    st, _, o_order, o_wire, o_gates, o_wc = orc.backend_raw(oc.gate_array(), oc.node_count + 1, nodes(ins), nodes(outs))
    assert st == 0
    info = ctx.emit_events(ev)
    g, _ = ctx.emitted_fetch(want_nodes=False)
    assert info["path"] == DEVICE and info["node_count"] == oc.node_count and np.array_equal(g, oc.gate_array())
    for implicit in (False, True):
        k, w, f = c2a.pack_events(ev, implicit=implicit)
        for fused in (0, 1 << 22):
            c2a.lib.c2a_set_fused_limits(fused, 0)
            try:
                i2, order, wire, ng, wc = ctx.compile_packed(k, w, f, ins, outs)
            finally:
                c2a.lib.c2a_set_fused_limits(1 << 22, 0)
            assert i2["node_count"] == oc.node_count and wc == o_wc
            assert np.array_equal(order, o_order) and np.array_equal(wire, o_wire) and np.array_equal(ng, o_gates)
