"""SHA-256 compression function, .circom text to digest, entirely through the product path: front end -> compressed recording ->
device emitter -> build -> named-wire lookup -> GPU evaluator; checked against hashlib."""
import hashlib
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sha256_from_circom_text_to_digest_on_the_gpu(c2a, ctx):
    dev = c2a.compile(None, source=c2a.workloads.sha256_circom_source(), context=ctx, emitter="device")
    info = ctx.emit_compressed(dev.compressed())
    assert info["path"] == 1 and info["n_gates"] == 3448
    _order, _wire, gates, wc = ctx.emitted_build_circuit(dev.input_signals, dev.output_signals, want_order=False, want_wires=False)
    in_names, out_names = dev.signal_names(dev.input_signals), dev.signal_names(dev.output_signals)
    in_w = ctx.emitted_signal_wires(dev.input_signals)
    out_w = ctx.emitted_signal_wires(dev.output_signals)
    const_w = ctx.emitted_signal_wires(dev._const_signals)
    iv = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]
    for msg in (b"abc", b"", b"gate graphs on B200"):
        blk = msg + b"\x80" + b"\0" * (55 - len(msg)) + struct.pack(">Q", 8 * len(msg))
        w = struct.unpack(">16I", blk)
        named = {f"0.h[{i}]": iv[i] for i in range(8)}
        named.update({f"0.w[{i}]": w[i] for i in range(16)})
        # "0.w" / "0.h" also tag the intermediate arrays ws[*] / hh[*] as inputs (prefix match): those wires are written by gates
        vals = {int(wi): named[n] for n, wi in zip(in_names, in_w) if n in named}
        vals.update({int(wi): int(v) for wi, v in zip(const_w, dev._const_values)})
        got = ctx.evaluate(gates, wc, vals)
        out = {n: int(got[int(wi)]) for n, wi in zip(out_names, out_w)}
        digest = b"".join(struct.pack(">I", out[f"0.out[{i}]"]) for i in range(8)).hex()
        assert digest == hashlib.sha256(msg).hexdigest()
