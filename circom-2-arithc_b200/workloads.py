"""Synthetic emission streams shaped like the BASELINE.json configs.

The reference cannot compile circomlib Poseidon / SHA-256 / Keccak / MiMC7 (u32-only constants,
src/process.rs:302; no `===`, :187; no inline arrays, :310), so these are event streams in the order the
reference's walker WOULD emit them for an equivalent template written in its supported subset:
  * a binary expression touching a signal emits: [const signal for a var operand], tmp signal, gate
    (src/process.rs:426-478, 538-579);  `lhs <== expr` then connects tmp -> lhs (:241-273)
  * constants are one signal per value per context; loop-body contexts are dropped per iteration
    (src/runtime.rs:151-187), template calls start from an empty context (:75-77)
  * a template body is emitted at the call site, BEFORE the caller wires the component inputs
    (src/process.rs:353-367 vs :218-236)
Signal ids are sequential (src/runtime.rs:120-125).
"""
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np

from .gate_types import AGateType as G, EV_CONNECT, EV_GATE, EV_SIGNAL, EV_SIGNAL_CONST


@dataclass
class Workload:
    name: str
    events: np.ndarray            # (n,4) u32: kind|op<<8, a, b, c   (c2a_event)
    inputs: Dict[int, str]        # main-template input signals  (id -> "0.<name>")
    outputs: Dict[int, str]
    n_gates: int
    n_signals: int
    meta: dict = field(default_factory=dict)


class Emitter:
    """Records the add_signal / add_gate / add_connection calls the walker makes."""

    def __init__(self, first_signal: int = 0):
        self.next = first_signal
        self.ev: List[tuple] = []
        self.names: Dict[int, str] = {}
        self._consts: List[Dict[int, int]] = [{}]
        self.n_gates = 0

    # contexts (only the const-signal cache matters for emission)
    def push_call(self):
        self._consts.append({})

    def push_scope(self):
        self._consts.append(dict(self._consts[-1]))

    def pop(self):
        self._consts.pop()

    def signal(self, name: Optional[str] = None) -> int:
        i = self.next
        self.next += 1
        self.ev.append((EV_SIGNAL, i, 0, 0))
        if name is not None:
            self.names[i] = name
        return i

    def signals(self, n: int) -> List[int]:
        return [self.signal() for _ in range(n)]

    def const(self, value: int) -> int:  # make_constant, src/process.rs:558-579
        c = self._consts[-1]
        if value in c:
            return c[value]
        i = self.next
        self.next += 1
        self.ev.append((EV_SIGNAL_CONST, i, value & 0xFFFFFFFF, 0))
        c[value] = i
        return i

    def gate(self, op: int, lhs: int, rhs: int) -> int:
        out = self.signal()
        self.ev.append((EV_GATE | (int(op) << 8), lhs, rhs, out))
        self.n_gates += 1
        return out

    def connect(self, a: int, b: int):
        self.ev.append((EV_CONNECT, a, b, 0))

    def assign(self, lhs_signal: int, op: int, a: int, b: int):  # lhs <== a op b
        self.connect(self.gate(op, a, b), lhs_signal)

    def assign_c(self, lhs_signal: int, op: int, a: int, value: int):  # lhs <== a op <var>
        c = self.const(value)
        self.connect(self.gate(op, a, c), lhs_signal)

    def array(self) -> np.ndarray:
        return np.asarray(self.ev, dtype=np.uint32).reshape(-1, 4)


def _tile(instance: Callable[[Emitter, int], None], base0: int, W: int):
    """Replicate a per-instance emission W times.  Instances are structurally identical and every operand is
    affine in the instance index w, so instance w = instance 0 + w * (instance 1 - instance 0); checked on w=2."""
    def run(w, first):
        em = Emitter(first)
        instance(em, w)
        return em
    e0 = run(0, base0)
    S = e0.next - base0
    a0 = e0.array().astype(np.int64)
    if W == 1:
        return a0.astype(np.uint32), S, e0.n_gates
    a1 = run(1, base0 + S).array().astype(np.int64)
    stride = a1 - a0
    assert (stride[:, 0] == 0).all(), "instances are not structurally identical"
    if W > 2:
        a2 = run(2, base0 + 2 * S).array().astype(np.int64)
        assert (a2 == a0 + 2 * stride).all(), "operands are not affine in the instance index"
    w = np.arange(W, dtype=np.int64)[:, None, None]
    ev = (a0[None] + stride[None] * w).reshape(-1, 4)
    return ev.astype(np.uint32), S, e0.n_gates


def mimc_chains(W: int, rounds: int = 91, variant: str = "inorder") -> Workload:
    """W independent MiMC-shaped chains under one main (config 5).  6 gates per round:
    a=x+k, b=a+c_i, t2=b*b, t4=t2*t2, t6=t4*t2, t7=t6*b  (c_i = i).  G = 6*rounds*W (+W for 'late').
    variant 'inorder': `c[w].x_in <== in[w]`          -> gate vector already in dependency order
            'late'   : `c[w].x_in <== in[w] + 1` emitted after the chain body -> non-identity DFS order."""
    assert variant in ("inorder", "late")
    n = rounds
    main = Emitter(0)
    ins = [main.signal(f"0.in[{w}]") for w in range(W)]
    k = main.signal("0.k")
    outs = [main.signal(f"0.out[{w}]") for w in range(W)]
    base0 = main.next

    def instance(em: Emitter, w: int):
        em.push_scope()      # main's `for` body context
        em.push_call()       # template call: fresh context
        x_in, k_c, out_c = em.signal(), em.signal(), em.signal()
        a, b, t2, t4, t6, t7 = (em.signals(n) for _ in range(6))
        for i in range(n):
            em.push_scope()
            x = x_in if i == 0 else t7[i - 1]
            em.assign(a[i], G.AAdd, x, k_c)
            em.assign_c(b[i], G.AAdd, a[i], i)
            em.assign(t2[i], G.AMul, b[i], b[i])
            em.assign(t4[i], G.AMul, t2[i], t2[i])
            em.assign(t6[i], G.AMul, t4[i], t2[i])
            em.assign(t7[i], G.AMul, t6[i], b[i])
            em.pop()
        em.connect(t7[n - 1], out_c)          # out <== t7[n-1]
        em.pop()
        if variant == "late":
            em.assign_c(x_in, G.AAdd, w, 1)   # c[w].x_in <== in[w] + 1   (in[w] has signal id w)
        else:
            em.connect(w, x_in)               # c[w].x_in <== in[w]
        em.connect(W, k_c)                    # c[w].k <== k
        em.connect(out_c, W + 1 + w)          # out[w] <== c[w].out
        em.pop()

    ev, S, gpc = _tile(instance, base0, W)
    events = np.concatenate([main.array(), ev], axis=0)
    return Workload(
        name=f"mimc_chains(W={W},rounds={rounds},{variant})", events=events,
        inputs={**{i: f"0.in[{i}]" for i in range(W)}, k: "0.k"},
        outputs={o: f"0.out[{o - W - 1}]" for o in outs},
        n_gates=gpc * W, n_signals=base0 + S * W, meta={"W": W, "rounds": rounds, "variant": variant, "signals_per_chain": S})


def mimc_circom_source(W: int, rounds: int = 91) -> str:
    """BASELINE config 5 as a real .circom program (W chains of `rounds` MiMC-7 rounds x -> (x + k + c_i)^7, c_i = i, u32 arithmetic):
    what the front end (csrc/c2a_front.cpp) walks when the pipeline is driven from source text instead of a synthetic stream."""
    return """pragma circom 2.0.0;
template Round(c) {
    signal input x; signal input k; signal output y;
    signal t; signal t2; signal t4; signal t6;
    t <== x + k + c;
    t2 <== t * t; t4 <== t2 * t2; t6 <== t4 * t2;
    y <== t6 * t;
}
template MiMC(n) {
    signal input x_in; signal input k; signal output out;
    component r[n];
    for (var i = 0; i < n; i++) {
        r[i] = Round(i);
        r[i].k <== k;
        if (i == 0) { r[i].x <== x_in; } else { r[i].x <== r[i - 1].y; }
    }
    out <== r[n - 1].y + k;
}
template Main(W, n) {
    signal input in[W]; signal input key; signal output out[W];
    component m[W];
    for (var w = 0; w < W; w++) { m[w] = MiMC(n); m[w].x_in <== in[w]; m[w].k <== key; out[w] <== m[w].out; }
}
component main = Main(%d, %d);
""" % (W, rounds)


def poseidon_shaped(t: int = 3, RF: int = 8, RP: int = 57) -> Workload:
    """Poseidon(2)-shaped permutation (config 2): ARC = AAdd with per-round constants c=(round*t+lane+1),
    S-box x^5 = 3 AMul, MDS = t*t AMul-by-const + t*(t-1) AAdd.  t=3,RF=8,RP=57 -> 1413 gates."""
    em = Emitter(0)
    ins = [em.signal(f"0.in[{i}]") for i in range(t - 1)]
    out = em.signal("0.out")
    state = [em.signal() for _ in range(t)]
    em.connect(em.const(0), state[0])  # capacity lane <== 0
    for i in range(t - 1):
        em.connect(ins[i], state[i + 1])
    nr = RF + RP
    for r in range(nr):
        em.push_scope()
        full = r < RF // 2 or r >= RF // 2 + RP
        nxt = []
        for lane in range(t):
            nxt.append(em.gate(G.AAdd, state[lane], em.const(r * t + lane + 1)))
        for lane in range(t if full else 1):
            x = nxt[lane]
            x2 = em.gate(G.AMul, x, x)
            x4 = em.gate(G.AMul, x2, x2)
            nxt[lane] = em.gate(G.AMul, x4, x)
        mixed = []
        for i in range(t):
            acc = None
            for j in range(t):
                term = em.gate(G.AMul, nxt[j], em.const(((i + 1) * (j + 2)) % 7 + 1))
                acc = term if acc is None else em.gate(G.AAdd, acc, term)
            s = em.signal()
            em.connect(acc, s)
            mixed.append(s)
        state = mixed
        em.pop()
    em.connect(state[0], out)
    return Workload(name=f"poseidon_shaped(t={t},RF={RF},RP={RP})", events=em.array(),
                    inputs={i: em.names[i] for i in ins}, outputs={out: "0.out"}, n_gates=em.n_gates, n_signals=em.next)


def poseidon_circom_source(t: int = 3, RF: int = 8, RP: int = 57) -> str:
    """BASELINE config 2 as a .circom program in the subset the reference accepts (circomlib's own Poseidon uses features it rejects):
    the same permutation as poseidon_shaped() - ARC constants c = round*t + lane + 1, S-box x^5 (3 multiplications), MDS entry
    ((i+1)*(j+2)) % 7 + 1 - with the capacity lane fed by 0 and lane 0 of the last state as output; u32 arithmetic (wrapping)."""
    return """pragma circom 2.0.0;
function arc(r, lane, t) { return r * t + lane + 1; }
function mds(i, j) { return ((i + 1) * (j + 2)) %% 7 + 1; }
template Sbox() {
    signal input in; signal output out;
    signal x2; signal x4;
    x2 <== in * in; x4 <== x2 * x2; out <== x4 * in;
}
template Round(r, full, t) {
    signal input in[t]; signal output out[t];
    signal a[t]; signal b[t];
    component s[t];
    for (var i = 0; i < t; i++) {
        a[i] <== in[i] + arc(r, i, t);
        if (full == 1 || i == 0) { s[i] = Sbox(); s[i].in <== a[i]; b[i] <== s[i].out; } else { b[i] <== a[i]; }
    }
    signal acc[t][t];
    for (var i = 0; i < t; i++) {
        acc[i][0] <== b[0] * mds(i, 0);
        for (var j = 1; j < t; j++) { acc[i][j] <== acc[i][j - 1] + b[j] * mds(i, j); }
        out[i] <== acc[i][t - 1];
    }
}
template Poseidon(t, RF, RP) {
    signal input in[t - 1]; signal output out;
    component rounds[RF + RP];
    for (var r = 0; r < RF + RP; r++) {
        var full = 0;
        if (r < RF \\ 2 || r >= RF \\ 2 + RP) { full = 1; }
        rounds[r] = Round(r, full, t);
        if (r == 0) {
            rounds[r].in[0] <== 0;
            for (var i = 1; i < t; i++) { rounds[r].in[i] <== in[i - 1]; }
        } else {
            for (var i = 0; i < t; i++) { rounds[r].in[i] <== rounds[r - 1].out[i]; }
        }
    }
    out <== rounds[RF + RP - 1].out[0];
}
component main = Poseidon(%d, %d, %d);
""" % (t, RF, RP)


def poseidon_reference(inputs: List[int], t: int = 3, RF: int = 8, RP: int = 57) -> int:
    """the function poseidon_circom_source() / poseidon_shaped() compute, in u32 arithmetic"""
    M = 0xFFFFFFFF
    state = [0] + [int(x) & M for x in inputs]
    for r in range(RF + RP):
        full = r < RF // 2 or r >= RF // 2 + RP
        a = [(state[i] + r * t + i + 1) & M for i in range(t)]
        b = [pow(a[i], 5, 1 << 32) if (full or i == 0) else a[i] for i in range(t)]
        state = [sum(b[j] * (((i + 1) * (j + 2)) % 7 + 1) for j in range(t)) & M for i in range(t)]
    return state[0]


_SHA256_K = [1116352408, 1899447441, 3049323471, 3921009573, 961987163, 1508970993, 2453635748, 2870763221, 3624381080, 310598401, 607225278, 1426881987, 1925078388, 2162078206, 2614888103, 3248222580, 3835390401, 4022224774, 264347078, 604807628, 770255983, 1249150122, 1555081692, 1996064986, 2554220882, 2821834349, 2952996808, 3210313671, 3336571891, 3584528711, 113926993, 338241895, 666307205, 773529912, 1294757372, 1396182291, 1695183700, 1986661051, 2177026350, 2456956037, 2730485921, 2820302411, 3259730800, 3345764771, 3516065817, 3600352804, 4094571909, 275423344, 430227734, 506948616, 659060556, 883997877, 958139571, 1322822218, 1537002063, 1747873779, 1955562222, 2024104815, 2227730452, 2361852424, 2428436474, 2756734187, 3204031479, 3329325298]


def sha256_circom_source() -> str:
    """The SHA-256 compression function (FIPS 180-4 section 6.2.2) as a .circom program in the subset the reference accepts, on the
    u32 arithmetic of its gates (wrapping +, >>, <<, &, ^, |, ~): inputs h[8] (chaining value) and w[16] (message block, big-endian
    words), outputs out[8].  ~3.6 K gates; BASELINE config 3 names the bit-sliced variant of the same function (sha256_shaped())."""
    return """pragma circom 2.0.0;
""" + 'function K(i) {\n    var r = 0;\n    if (i == 0) { r = 0x428a2f98; }\n    if (i == 1) { r = 0x71374491; }\n    if (i == 2) { r = 0xb5c0fbcf; }\n    if (i == 3) { r = 0xe9b5dba5; }\n    if (i == 4) { r = 0x3956c25b; }\n    if (i == 5) { r = 0x59f111f1; }\n    if (i == 6) { r = 0x923f82a4; }\n    if (i == 7) { r = 0xab1c5ed5; }\n    if (i == 8) { r = 0xd807aa98; }\n    if (i == 9) { r = 0x12835b01; }\n    if (i == 10) { r = 0x243185be; }\n    if (i == 11) { r = 0x550c7dc3; }\n    if (i == 12) { r = 0x72be5d74; }\n    if (i == 13) { r = 0x80deb1fe; }\n    if (i == 14) { r = 0x9bdc06a7; }\n    if (i == 15) { r = 0xc19bf174; }\n    if (i == 16) { r = 0xe49b69c1; }\n    if (i == 17) { r = 0xefbe4786; }\n    if (i == 18) { r = 0x0fc19dc6; }\n    if (i == 19) { r = 0x240ca1cc; }\n    if (i == 20) { r = 0x2de92c6f; }\n    if (i == 21) { r = 0x4a7484aa; }\n    if (i == 22) { r = 0x5cb0a9dc; }\n    if (i == 23) { r = 0x76f988da; }\n    if (i == 24) { r = 0x983e5152; }\n    if (i == 25) { r = 0xa831c66d; }\n    if (i == 26) { r = 0xb00327c8; }\n    if (i == 27) { r = 0xbf597fc7; }\n    if (i == 28) { r = 0xc6e00bf3; }\n    if (i == 29) { r = 0xd5a79147; }\n    if (i == 30) { r = 0x06ca6351; }\n    if (i == 31) { r = 0x14292967; }\n    if (i == 32) { r = 0x27b70a85; }\n    if (i == 33) { r = 0x2e1b2138; }\n    if (i == 34) { r = 0x4d2c6dfc; }\n    if (i == 35) { r = 0x53380d13; }\n    if (i == 36) { r = 0x650a7354; }\n    if (i == 37) { r = 0x766a0abb; }\n    if (i == 38) { r = 0x81c2c92e; }\n    if (i == 39) { r = 0x92722c85; }\n    if (i == 40) { r = 0xa2bfe8a1; }\n    if (i == 41) { r = 0xa81a664b; }\n    if (i == 42) { r = 0xc24b8b70; }\n    if (i == 43) { r = 0xc76c51a3; }\n    if (i == 44) { r = 0xd192e819; }\n    if (i == 45) { r = 0xd6990624; }\n    if (i == 46) { r = 0xf40e3585; }\n    if (i == 47) { r = 0x106aa070; }\n    if (i == 48) { r = 0x19a4c116; }\n    if (i == 49) { r = 0x1e376c08; }\n    if (i == 50) { r = 0x2748774c; }\n    if (i == 51) { r = 0x34b0bcb5; }\n    if (i == 52) { r = 0x391c0cb3; }\n    if (i == 53) { r = 0x4ed8aa4a; }\n    if (i == 54) { r = 0x5b9cca4f; }\n    if (i == 55) { r = 0x682e6ff3; }\n    if (i == 56) { r = 0x748f82ee; }\n    if (i == 57) { r = 0x78a5636f; }\n    if (i == 58) { r = 0x84c87814; }\n    if (i == 59) { r = 0x8cc70208; }\n    if (i == 60) { r = 0x90befffa; }\n    if (i == 61) { r = 0xa4506ceb; }\n    if (i == 62) { r = 0xbef9a3f7; }\n    if (i == 63) { r = 0xc67178f2; }\n    return r;\n}\n' + """template Rotr(n) { signal input x; signal output y; y <== (x >> n) | (x << (32 - n)); }
template BigSigma(a, b, c) {
    signal input x; signal output y;
    component r0 = Rotr(a); component r1 = Rotr(b); component r2 = Rotr(c);
    r0.x <== x; r1.x <== x; r2.x <== x;
    y <== r0.y ^ r1.y ^ r2.y;
}
template SmallSigma(a, b, s) {
    signal input x; signal output y;
    component r0 = Rotr(a); component r1 = Rotr(b);
    r0.x <== x; r1.x <== x;
    y <== r0.y ^ r1.y ^ (x >> s);
}
template Ch() { signal input e; signal input f; signal input g; signal output y; y <== (e & f) ^ ((~e) & g); }
template Maj() { signal input a; signal input b; signal input c; signal output y; y <== (a & b) ^ (a & c) ^ (b & c); }
template Sha256Compress() {
    signal input h[8]; signal input w[16]; signal output out[8];
    signal ws[64];
    component s0[64]; component s1[64];
    for (var i = 0; i < 64; i++) {
        if (i < 16) { ws[i] <== w[i]; } else {
            s0[i] = SmallSigma(7, 18, 3); s1[i] = SmallSigma(17, 19, 10);
            s0[i].x <== ws[i - 15]; s1[i].x <== ws[i - 2];
            ws[i] <== s1[i].y + ws[i - 7] + s0[i].y + ws[i - 16];
        }
    }
    signal a[65]; signal b[65]; signal c[65]; signal d[65]; signal e[65]; signal f[65]; signal g[65]; signal hh[65];
    a[0] <== h[0]; b[0] <== h[1]; c[0] <== h[2]; d[0] <== h[3]; e[0] <== h[4]; f[0] <== h[5]; g[0] <== h[6]; hh[0] <== h[7];
    component S0[64]; component S1[64]; component ch[64]; component maj[64];
    signal t1[64]; signal t2[64];
    for (var i = 0; i < 64; i++) {
        S1[i] = BigSigma(6, 11, 25); S1[i].x <== e[i];
        ch[i] = Ch(); ch[i].e <== e[i]; ch[i].f <== f[i]; ch[i].g <== g[i];
        t1[i] <== hh[i] + S1[i].y + ch[i].y + K(i) + ws[i];
        S0[i] = BigSigma(2, 13, 22); S0[i].x <== a[i];
        maj[i] = Maj(); maj[i].a <== a[i]; maj[i].b <== b[i]; maj[i].c <== c[i];
        t2[i] <== S0[i].y + maj[i].y;
        hh[i + 1] <== g[i]; g[i + 1] <== f[i]; f[i + 1] <== e[i]; e[i + 1] <== d[i] + t1[i];
        d[i + 1] <== c[i]; c[i + 1] <== b[i]; b[i + 1] <== a[i]; a[i + 1] <== t1[i] + t2[i];
    }
    out[0] <== h[0] + a[64]; out[1] <== h[1] + b[64]; out[2] <== h[2] + c[64]; out[3] <== h[3] + d[64];
    out[4] <== h[4] + e[64]; out[5] <== h[5] + f[64]; out[6] <== h[6] + g[64]; out[7] <== h[7] + hh[64];
}
component main = Sha256Compress();
"""


def sha256_compress_reference(h: List[int], w: List[int]) -> List[int]:
    """FIPS 180-4 section 6.2.2 on Python integers"""
    M = 0xFFFFFFFF
    rotr = lambda x, n: ((x >> n) | (x << (32 - n))) & M
    ws = [int(x) & M for x in w]
    for i in range(16, 64):
        s0 = rotr(ws[i - 15], 7) ^ rotr(ws[i - 15], 18) ^ (ws[i - 15] >> 3)
        s1 = rotr(ws[i - 2], 17) ^ rotr(ws[i - 2], 19) ^ (ws[i - 2] >> 10)
        ws.append((s1 + ws[i - 7] + s0 + ws[i - 16]) & M)
    a, b, c, d, e, f, g, hh = [int(x) & M for x in h]
    for i in range(64):
        t1 = (hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & M & g)) + _SHA256_K[i] + ws[i]) & M
        t2 = ((rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c))) & M
        a, b, c, d, e, f, g, hh = (t1 + t2) & M, a, b, c, (d + t1) & M, e, f, g
    return [(x + y) & M for x, y in zip([int(x) & M for x in h], [a, b, c, d, e, f, g, hh])]


def _xor(em, a, b):
    return em.gate(G.AXor, a, b)


def _and(em, a, b):
    return em.gate(G.ABitAnd, a, b)


def _add32(em: Emitter, a: List[int], b: List[int]) -> List[int]:
    """bit-sliced ripple-carry adder, 5 gates per bit: s=a^b^c; c'=(a&b)^(c&(a^b))"""
    out, carry = [], None
    for i in range(len(a)):
        axb = _xor(em, a[i], b[i])
        if carry is None:
            out.append(axb)
            carry = _and(em, a[i], b[i])
        else:
            out.append(_xor(em, axb, carry))
            carry = _xor(em, _and(em, a[i], b[i]), _and(em, carry, axb))
    return out


def _rot(a, r):
    return a[r:] + a[:r]


def sha256_shaped(rounds: int = 64) -> Workload:
    """SHA-256-compression-shaped bit-sliced circuit (config 3): Σ/σ/Ch/Maj as AXor/ABitAnd on single-bit
    signals, 32-bit ripple-carry adders. ~30 K gates at 64 rounds, depth in the thousands."""
    em = Emitter(0)
    w_in = [[em.signal(f"0.w[{i}][{b}]") for b in range(32)] for i in range(16)]
    outs = [[em.signal(f"0.h[{i}][{b}]") for b in range(32)] for i in range(8)]
    st = [[em.const((0x6A09E667 >> b) & 1 if i % 2 == 0 else (0xBB67AE85 >> b) & 1) for b in range(32)] for i in range(8)]
    w = list(w_in)
    for r in range(rounds):
        em.push_scope()
        a, b, c, d, e, f, g, h = st
        if r >= 16:
            x, y = w[r - 15], w[r - 2]
            s0 = [_xor(em, _xor(em, p, q), z) for p, q, z in zip(_rot(x, 7), _rot(x, 18), _rot(x, 3))]
            s1 = [_xor(em, _xor(em, p, q), z) for p, q, z in zip(_rot(y, 17), _rot(y, 19), _rot(y, 10))]
            w.append(_add32(em, _add32(em, w[r - 16], s0), _add32(em, w[r - 7], s1)))
        S1 = [_xor(em, _xor(em, p, q), z) for p, q, z in zip(_rot(e, 6), _rot(e, 11), _rot(e, 25))]
        ch = [_xor(em, _and(em, p, q), _and(em, _xor(em, p, em.const(1)), z)) for p, q, z in zip(e, f, g)]
        t1 = _add32(em, _add32(em, h, S1), _add32(em, ch, w[r]))
        S0 = [_xor(em, _xor(em, p, q), z) for p, q, z in zip(_rot(a, 2), _rot(a, 13), _rot(a, 22))]
        mj = [_xor(em, _xor(em, _and(em, p, q), _and(em, p, z)), _and(em, q, z)) for p, q, z in zip(a, b, c)]
        t2 = _add32(em, S0, mj)
        st = [_add32(em, t1, t2), a, b, c, _add32(em, d, t1), e, f, g]
        em.pop()
    for i in range(8):
        for b_ in range(32):
            # out <== state bit XOR 0-extended: keep outputs produced by gates (outputs cannot merge with consts)
            em.connect(_xor(em, st[i][b_], w_in[i][b_]), outs[i][b_])
    ins = {s: em.names[s] for row in w_in for s in row}
    return Workload(name=f"sha256_shaped(rounds={rounds})", events=em.array(), inputs=ins,
                    outputs={s: em.names[s] for row in outs for s in row}, n_gates=em.n_gates, n_signals=em.next)


def keccak_shaped(instances: int = 1, rounds: int = 24) -> Workload:
    """Keccak-f[1600]-shaped bit-sliced sponge permutation (config 4): θ (column parities, XOR), ρ/π as pure
    rewiring, χ (a ^ (~b & c)), ι (XOR with a round constant on lane 0).  ~150 K gates per instance at 24 rounds.
    `instances` independent permutations under one main = independent components for the 2-GPU shard."""
    main = Emitter(0)
    nin = 1600
    ins = [[main.signal(f"0.in[{k}][{i}]") for i in range(nin)] for k in range(instances)]
    outs = [[main.signal(f"0.out[{k}][{i}]") for i in range(nin)] for k in range(instances)]
    base0 = main.next
    rho = [0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14]

    def instance(em: Emitter, k: int):
        em.push_scope()
        em.push_call()
        s_in = [em.signal() for _ in range(nin)]
        s_out = [em.signal() for _ in range(nin)]
        A = [[s_in[(x + 5 * y) * 64:(x + 5 * y) * 64 + 64] for y in range(5)] for x in range(5)]
        for r in range(rounds):
            em.push_scope()
            Cp = [[_xor(em, _xor(em, _xor(em, _xor(em, A[x][0][z], A[x][1][z]), A[x][2][z]), A[x][3][z]), A[x][4][z]) for z in range(64)] for x in range(5)]
            D = [[_xor(em, Cp[(x - 1) % 5][z], Cp[(x + 1) % 5][(z - 1) % 64]) for z in range(64)] for x in range(5)]
            A = [[[_xor(em, A[x][y][z], D[x][z]) for z in range(64)] for y in range(5)] for x in range(5)]
            B = [[None] * 5 for _ in range(5)]
            for x in range(5):
                for y in range(5):
                    rr = rho[x + 5 * y] % 64
                    B[y][(2 * x + 3 * y) % 5] = A[x][y][-rr:] + A[x][y][:-rr] if rr else A[x][y]
            one = em.const(1)
            A = [[[_xor(em, B[x][y][z], _and(em, _xor(em, B[(x + 1) % 5][y][z], one), B[(x + 2) % 5][y][z])) for z in range(64)] for y in range(5)] for x in range(5)]
            rc = (0x8000000080008081 * (r + 1) + 0x9E3779B97F4A7C15 * r) & ((1 << 64) - 1)
            A[0][0] = [_xor(em, A[0][0][z], em.const((rc >> z) & 1)) for z in range(64)]
            em.pop()
        for x in range(5):
            for y in range(5):
                for z in range(64):
                    em.connect(A[x][y][z], s_out[(x + 5 * y) * 64 + z])
        em.pop()
        # caller wiring AFTER the body (src/process.rs:353-367 vs :218-236): in -> component, component -> out
        for i in range(nin):
            em.connect(k * nin + i, s_in[i])
        for i in range(nin):
            em.connect(s_out[i], instances * nin + k * nin + i)
        em.pop()

    ev, S, gpc = _tile(instance, base0, instances)
    events = np.concatenate([main.array(), ev], axis=0)
    return Workload(name=f"keccak_shaped(instances={instances},rounds={rounds})", events=events,
                    inputs={s: main.names[s] for row in ins for s in row}, outputs={s: main.names[s] for row in outs for s in row},
                    n_gates=gpc * instances, n_signals=base0 + S * instances, meta={"instances": instances})


def shuffle_gates(gates: np.ndarray, seed: int = 1) -> np.ndarray:
    """Stress variant: permute the gate VECTOR (node ids unchanged) -> heavily out-of-order DFS roots."""
    rng = np.random.RandomState(seed)
    return np.ascontiguousarray(gates[rng.permutation(gates.shape[0])])


def random_gates(G: int, n_free: int, seed: int, p_forward: float = 0.2, p_dup_out: float = 0.02, window: int = 0) -> tuple:
    """Random gate vector in NODE-id form for back-end parity tests: node ids 1..n_free are producer-less
    (inputs/constants), gate g writes node n_free+1+perm[g] (so producers are scattered and a fraction of
    operands refer to gates that come LATER in the vector).  Cycle-free by construction: operands are chosen
    among nodes whose producing gate has a smaller rank in a hidden topological order.
    Returns (gates (G,4) u32, node_bound)."""
    rng = np.random.RandomState(seed)
    rank_of_gate = rng.permutation(G) if p_forward > 0 else np.arange(G)   # hidden topological rank
    if p_forward < 1.0 and p_forward > 0:
        keep = rng.rand(G) >= p_forward
        rank_of_gate = np.where(keep, np.arange(G), rank_of_gate)
        rank_of_gate = np.argsort(np.argsort(rank_of_gate, kind="stable"), kind="stable")
    gate_of_rank = np.argsort(rank_of_gate)
    out_node = n_free + 1 + np.arange(G)
    gates = np.zeros((G, 4), dtype=np.uint32)
    gates[:, 0] = rng.randint(0, 20, size=G)
    gates[:, 3] = out_node
    for slot in (1, 2):
        rk = rank_of_gate
        lo = np.maximum(0, rk - window) if window else np.zeros(G, dtype=np.int64)
        pick_rank = (lo + (rng.rand(G) * np.maximum(rk - lo, 0)).astype(np.int64))
        use_free = (rng.rand(G) < 0.3) | (rk == 0)
        src_gate = gate_of_rank[np.minimum(pick_rank, np.maximum(rk - 1, 0))]
        gates[:, slot] = np.where(use_free, rng.randint(1, n_free + 1, size=G), out_node[src_gate])
    if p_dup_out > 0 and G > 4:
        # a few gates share an out node with an EARLIER-ranked gate (merged nodes): last writer wins upstream
        dup = np.where(rng.rand(G) < p_dup_out)[0]
        for g in dup:
            r = rank_of_gate[g]
            if r > 0:
                gates[g, 3] = out_node[gate_of_rank[rng.randint(0, r)]]
    return gates, int(n_free + 1 + G)
