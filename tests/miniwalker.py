"""A tiny stand-in for src/process.rs used by the tests until .circom fixtures can be parsed: it issues the same
add_signal / add_gate / add_connection calls, in the same order and with the same names, as the reference's
walker does for the handful of statement shapes the integration fixtures use.
Works on any object with the Compiler interface (oracle or product)."""


class MiniWalker:
    def __init__(self, comp, ctx="0"):
        self.c = comp
        self.ctx = ctx
        self.next = 0
        self.consts = {}

    def _new(self):
        i = self.next
        self.next += 1
        return i

    def signal(self, name):  # Declaration of a scalar signal (src/process.rs:79-88)
        i = self._new()
        self.c.add_signal(i, f"{self.ctx}.{name}", None)
        return i

    def array(self, name, *dims):  # row-major (src/runtime.rs:431-445, src/process.rs:89-108)
        import itertools
        ids = {}
        for idx in itertools.product(*[range(d) for d in dims]):
            i = self._new()
            self.c.add_signal(i, f"{self.ctx}.{name}" + "".join(f"[{k}]" for k in idx), None)
            ids[idx] = i
        return ids

    def const(self, value):  # make_constant (src/process.rs:558-579)
        if value not in self.consts:
            i = self._new()
            self.c.add_signal(i, f"{self.ctx}.const_signal_{value}", value)
            self.consts[value] = i
        return self.consts[value]

    def infix(self, op, lhs, rhs):
        """operands: int signal id, or ('var', value). Returns the tmp output signal (src/process.rs:426-478)."""
        l = self.const(lhs[1]) if isinstance(lhs, tuple) else lhs
        r = self.const(rhs[1]) if isinstance(rhs, tuple) else rhs
        out = self._new()
        self.c.add_signal(out, f"{self.ctx}.random_{out}", None)
        self.c.add_gate(op, l, r, out)
        return out

    def prefix(self, kind, rhs):  # src/process.rs:485-533, 758-764
        value, op = {"neg": (0, 9), "not": (0, 2), "complement": (0xFFFFFFFF, 10)}[kind]
        l = self.const(value)
        out = self._new()
        self.c.add_signal(out, f"{self.ctx}.random_{out}", None)
        self.c.add_gate(op, l, rhs, out)
        return out

    def assign(self, lhs_signal, rhs):  # lhs <== <signal | ('var', v)>  (src/process.rs:241-273)
        src = self.const(rhs[1]) if isinstance(rhs, tuple) else rhs
        self.c.add_connection(src, lhs_signal)

    def tag_io(self, inputs, outputs):  # src/program.rs:57-66 (prefix match!)
        for n in inputs:
            self.c.tag_inputs_by_prefix(f"0.{n}")
        for n in outputs:
            self.c.tag_outputs_by_prefix(f"0.{n}")


# op numbers = AGateType discriminants (src/a_gate_type.rs:7-28)
AAdd, ADiv, AEq, AGEq, AGt, ALEq, ALt, AMul, ANeq, ASub, AXor, APow, AIntDiv, AMod, AShiftL, AShiftR, ABoolOr, ABoolAnd, ABitOr, ABitAnd = range(20)


def fixture_add_zero(c):      # tests/circuits/integration/addZero.circom
    w = MiniWalker(c)
    i, o = w.signal("in"), w.signal("out")
    w.assign(o, w.infix(AAdd, i, ("var", 0)))
    w.tag_io(["in"], ["out"])


def fixture_sum(c):           # sum.circom
    w = MiniWalker(c)
    a, b, o = w.signal("a"), w.signal("b"), w.signal("out")
    w.assign(o, w.infix(AAdd, a, b))
    w.tag_io(["a", "b"], ["out"])


def fixture_x_eq_x(c):        # xEqX.circom
    w = MiniWalker(c)
    x, o = w.signal("x"), w.signal("out")
    w.assign(o, w.infix(AEq, x, x))
    w.tag_io(["x"], ["out"])


def fixture_constant_sum(c):  # constantSum.circom: 3 + 5 folded on the host (src/process.rs:444-457)
    w = MiniWalker(c)
    o = w.signal("out")
    w.assign(o, ("var", 8))
    w.tag_io([], ["out"])


def fixture_direct_output(c):  # directOutput.circom
    w = MiniWalker(c)
    o = w.signal("out")
    w.assign(o, ("var", 42))
    w.tag_io([], ["out"])


INFIX_OUTPUTS = [  # infixOps.circom, in declaration order: (name, op, lhs input index, rhs input index, expected)
    ("mul_2_3", AMul, 2, 3, 6), ("idiv_4_3", AIntDiv, 4, 3, 1), ("add_3_4", AAdd, 3, 4, 7), ("sub_4_1", ASub, 4, 1, 3),
    ("pow_2_4", APow, 2, 4, 16), ("mod_5_3", AMod, 5, 3, 2), ("shl_5_1", AShiftL, 5, 1, 10), ("shr_5_1", AShiftR, 5, 1, 2),
    ("leq_2_3", ALEq, 2, 3, 1), ("leq_3_3", ALEq, 3, 3, 1), ("leq_4_3", ALEq, 4, 3, 0), ("geq_2_3", AGEq, 2, 3, 0),
    ("geq_3_3", AGEq, 3, 3, 1), ("geq_4_3", AGEq, 4, 3, 1), ("lt_2_3", ALt, 2, 3, 1), ("lt_3_3", ALt, 3, 3, 0),
    ("lt_4_3", ALt, 4, 3, 0), ("gt_2_3", AGt, 2, 3, 0), ("gt_3_3", AGt, 3, 3, 0), ("gt_4_3", AGt, 4, 3, 1),
    ("eq_2_3", AEq, 2, 3, 0), ("eq_3_3", AEq, 3, 3, 1), ("neq_2_3", ANeq, 2, 3, 1), ("neq_3_3", ANeq, 3, 3, 0),
    ("or_0_1", ABoolOr, 0, 1, 1), ("and_0_1", ABoolAnd, 0, 1, 0), ("bit_or_1_3", ABitOr, 1, 3, 3),
    ("bit_and_1_3", ABitAnd, 1, 3, 1), ("bit_xor_1_3", AXor, 1, 3, 2)]


def fixture_infix_ops(c):
    w = MiniWalker(c)
    x = [w.signal(f"x{i}") for i in range(6)]
    outs = [w.signal(n) for n, *_ in INFIX_OUTPUTS]
    for o, (_, op, l, r, _e) in zip(outs, INFIX_OUTPUTS):
        w.assign(o, w.infix(op, x[l], x[r]))
    w.tag_io([f"x{i}" for i in range(6)], [n for n, *_ in INFIX_OUTPUTS])


def fixture_mat_elem_mul(c, m=2, n=2):  # matElemMul.circom (loop contexts do not change emission here: no constants)
    w = MiniWalker(c)
    a, b, o = w.array("a", m, n), w.array("b", m, n), w.array("out", m, n)
    for i in range(m):
        for j in range(n):
            w.assign(o[i, j], w.infix(AMul, a[i, j], b[i, j]))
    w.tag_io(["a", "b"], ["out"])


def fixture_prefix_ops(c):  # prefixOps.circom (ignored upstream: prefix-match I/O bug)
    w = MiniWalker(c)
    a, b, cc = w.signal("a"), w.signal("b"), w.signal("c")
    names = ["negateA", "notA", "notB", "notC", "complementA", "complementB", "complementC"]
    o = [w.signal(n) for n in names]
    for dst, (kind, src) in zip(o, [("neg", a), ("not", a), ("not", b), ("not", cc), ("complement", a), ("complement", b), ("complement", cc)]):
        w.assign(dst, w.prefix(kind, src))
    w.tag_io(["a", "b", "c"], names)


def fixture_array_assignment(c):  # arrayAssignment.circom: callee body first, caller wiring after
    main = MiniWalker(c, "0")
    a_in = main.array("a_in", 2, 2)
    out = main.signal("out")
    callee = MiniWalker(c, "componentA")
    callee.next = main.next
    cin = callee.array("in", 2, 2)
    cout = callee.signal("out")
    t = callee.infix(AAdd, cin[0, 0], cin[0, 1])
    t = callee.infix(AAdd, t, cin[1, 0])
    t = callee.infix(AAdd, t, cin[1, 1])
    callee.assign(cout, t)
    main.next = callee.next
    for k in [(0, 0), (0, 1), (1, 0), (1, 1)]:   # a.in <== a_in  (connect_signal_arrays(component, assigned))
        c.add_connection(cin[k], a_in[k])
    main.assign(out, cout)
    main.tag_io(["a_in"], ["out"])
