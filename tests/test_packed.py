"""Packed event stream (include/c2a.h: kinds byte + payload words): host-side pack / unpack round trips.  No GPU."""
import numpy as np
import pytest

EV_S, EV_SC, EV_G, EV_C = 0, 1, 2, 3


def _zero_values(ev):
    ev = ev.copy()
    ev[(ev[:, 0] & 0xFF) == EV_SC, 2] = 0  # constant values are not part of the packed stream
    return ev


@pytest.mark.parametrize("variant", ["inorder", "late"])
def test_round_trip_workloads(c2a, variant):
    wl = c2a.workloads.mimc_chains(9, rounds=7, variant=variant)
    ev = np.ascontiguousarray(wl.events)
    kinds, words, flags = c2a.pack_events(ev)
    k = ev[:, 0] & 0xFF
    assert flags == 1                                   # workload ids are dense, like Runtime::gen_signal (src/runtime.rs:120-125)
    assert len(words) == 3 * (k == EV_G).sum() + 2 * (k == EV_C).sum()
    assert kinds.nbytes + words.nbytes < 0.45 * ev.nbytes
    assert np.array_equal(c2a.unpack_events(kinds, words, flags), _zero_values(ev))


def test_round_trip_explicit_ids_and_all_ops(c2a):
    rng = np.random.RandomState(3)
    ids = rng.permutation(500).astype(np.uint32) * 3 + 1
    ev = [(EV_SC, int(s), int(i), 0) if i % 7 == 0 else (EV_S, int(s), 0, 0) for i, s in enumerate(ids)]
    for op in range(20):
        a, b, o = (int(x) for x in rng.choice(ids, 3))
        ev.append((EV_G | (op << 8), a, b, o))
        ev.append((EV_C, a, o, 0))
    ev = np.asarray(ev, dtype=np.uint32)
    ev = ev[rng.permutation(len(ev))]                    # order is irrelevant to the format
    kinds, words, flags = c2a.pack_events(ev)
    assert flags == 0 and len(words) == 500 + 3 * 20 + 2 * 20
    assert np.array_equal(c2a.unpack_events(kinds, words, flags), _zero_values(ev))
    with pytest.raises(c2a.C2AError):
        c2a.unpack_events(kinds, words[:-1], flags)


def test_empty(c2a):
    kinds, words, flags = c2a.pack_events(np.zeros((0, 4), dtype=np.uint32))
    assert len(kinds) == 0 and len(words) == 0 and flags == 1
    assert c2a.unpack_events(kinds, words, flags).shape == (0, 4)


def test_program_packed_matches_events(c2a):
    """the front end hands out the recorded calls in both forms (c2a_program_events / c2a_program_packed)"""
    import ctypes as C
    from circom_2_arithc_b200._lib import lib, PackedEvents
    src = b"pragma circom 2.0.0; template T(){ signal input a; signal input b; signal output c; c <== a*b + 3; } component main = T();"
    p = lib.c2a_program_new()
    try:
        assert lib.c2a_program_compile_source(p, src, None, None) == 0, lib.c2a_program_error(p)
        n = lib.c2a_program_num_events(p)
        ev = np.ctypeslib.as_array(C.cast(lib.c2a_program_events(p), C.POINTER(C.c_uint32)), shape=(n, 4)).copy()
        pk = PackedEvents()
        assert lib.c2a_program_packed(p, C.byref(pk)) == 0 and pk.n_events == n
        kinds = np.ctypeslib.as_array(C.cast(pk.kinds, C.POINTER(C.c_uint8)), shape=(n,)).copy()
        words = np.ctypeslib.as_array(C.cast(pk.words, C.POINTER(C.c_uint32)), shape=(pk.n_words,)).copy()
        assert np.array_equal(c2a.unpack_events(kinds, words, pk.flags), _zero_values(ev))
        k2, w2, f2 = c2a.pack_events(ev)
        assert np.array_equal(k2, kinds) and np.array_equal(w2, words) and f2 == pk.flags
    finally:
        lib.c2a_program_free(p)


def test_device_compiler_records_without_a_gpu(c2a):
    """compile(emitter="device") needs no device until build_circuit(): it holds the recorded calls (both forms), the tagged
    inputs / outputs and the names"""
    src = "template T(){ signal input a; signal input b; signal output c; c <== a*b + 3; } component main = T();"
    dc = c2a.compile(None, source=src, emitter="device")
    host = c2a.compile(None, source=src) if c2a.have_device() else None
    assert isinstance(dc, c2a.DeviceCompiler) and dc._flags == 1
    assert np.array_equal(c2a.unpack_events(dc._kinds, dc._words, dc._flags), _zero_values(dc.events))
    assert [dc.signal_name(s) for s in dc.input_signals] == ["0.a", "0.b"]
    # the reference tags outputs by PREFIX match on "0.c" (src/program.rs:62-66), which also catches the constant's name
    assert [dc.signal_name(s) for s in dc.output_signals] == ["0.c", "0.const_signal_3"]
    if host is not None:
        assert np.array_equal(host.events, dc.events)


@pytest.mark.parametrize("seed", range(6))
def test_implicit_operands_round_trip(c2a, seed):
    """C2A_PACKED_IMPLICIT_OPERANDS: a gate whose out signal / a connection whose first signal is the signal declared last carries no
    word for it; mixed with events that do not qualify, invalid ops and kinds, references before any declaration"""
    rng = np.random.RandomState(50 + seed)
    ev, ns = [], 0
    for _ in range(int(rng.randint(1, 400))):
        x = rng.rand()
        if x < 0.4 or ns == 0:
            ev.append((EV_SC, ns, int(rng.randint(0, 9)), 0) if rng.rand() < 0.2 else (EV_S, ns, 0, 0))
            ns += 1
        elif x < 0.7:
            op = int(rng.randint(0, 20)) if rng.rand() < 0.95 else int(rng.choice([20, 31, 40, 63, 200]))
            out = ns - 1 if rng.rand() < 0.7 else int(rng.randint(0, ns + 2))
            ev.append((EV_G | (op << 8), int(rng.randint(0, ns + 2)), int(rng.randint(0, ns + 2)), out))
        else:
            a = ns - 1 if rng.rand() < 0.7 else int(rng.randint(0, ns + 2))
            ev.append((EV_C, a, int(rng.randint(0, ns + 2)), 0))
    if seed == 0:
        ev = [(EV_G, 0, 0, 0), (EV_C, 0, 0, 0)] + ev      # events before any declaration cannot use the implicit form
    ev = np.asarray(ev, dtype=np.uint32).reshape(-1, 4)
    k = ev[:, 0] & 0xFF
    if seed == 0:
        ev[k <= 1, 1] = np.arange((k <= 1).sum())
    kinds, words, flags = c2a.pack_events(ev, implicit=True)
    assert flags == 3
    k0, w0, f0 = c2a.pack_events(ev)
    assert f0 == 1 and len(words) <= len(w0) and np.array_equal(kinds & 3, k0 & 3)
    n_flagged = int(((kinds & 0x80) != 0).sum())
    assert len(words) == len(w0) - n_flagged
    back = c2a.unpack_events(kinds, words, flags)
    want = _zero_values(ev)
    bad_op = (k == EV_G) & ((ev[:, 0] >> 8) >= 31)        # out-of-range ops stay out of range, not bit-identical
    assert np.array_equal(back[~bad_op], want[~bad_op])
    assert ((back[bad_op, 0] >> 8) >= 20).all() and np.array_equal(back[bad_op, 1:], want[bad_op, 1:])


def test_implicit_operands_shrink_a_walker_stream(c2a):
    wl = c2a.workloads.mimc_chains(5, rounds=9, variant="late")
    ev = np.ascontiguousarray(wl.events)
    kinds, words, flags = c2a.pack_events(ev, implicit=True)
    k = ev[:, 0] & 0xFF
    G, Cn = int((k == EV_G).sum()), int((k == EV_C).sum())
    assert flags == 3 and len(words) <= 2 * G + 2 * Cn - G + 8       # every gate, and the connection after every gate, lost a word
    assert (kinds.nbytes + words.nbytes) / len(ev) < 4.2
    assert np.array_equal(c2a.unpack_events(kinds, words, flags), _zero_values(ev))
