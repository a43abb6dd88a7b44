"""The reference's integration fixtures (tests/circuits/integration/*.circom, wired to tests/integration.rs:279-475),
restated as strings so that the tests do not read /root/reference at run time.  Signal / template names are the contract
(they appear in the expected maps); layout and comments are ours."""

ADD_ZERO = """pragma circom 2.1.0;
template addZero() { signal input in; signal output out; out <== in + 0; }
component main = addZero();
"""

SUM = """pragma circom 2.1.0;
template sum() {
    signal input a; signal input b;
    signal output out;
    out <== a + b;
}
component main = sum();
"""

X_EQ_X = """pragma circom 2.1.0;
template xEqX() { signal input x; signal output out; out <== x == x; }
component main = xEqX();
"""

CONSTANT_SUM = """pragma circom 2.1.0;
template constantSum() { signal output out; out <== 3 + 5; }
component main = constantSum();
"""

DIRECT_OUTPUT = """pragma circom 2.1.0;
template directOutput() { signal output out; out <== 42; }
component main = directOutput();
"""

INDEX_OUT_OF_BOUNDS = """pragma circom 2.1.0;
template indexOutOfBounds() {
    signal arr[10];
    for (var i = 0; i < 100; i++) { arr[i] <== 1; }
}
component main = indexOutOfBounds();
"""

MAIN_TEMPLATE_ARGUMENT = """pragma circom 2.1.0;
template mainComponent(argument) { signal input in; signal output out; out <== in + argument; }
component main = mainComponent(100);
"""

MAT_ELEM_MUL = """pragma circom 2.1.0;
template matElemMul(m, n) {
    signal input a[m][n];
    signal input b[m][n];
    signal output out[m][n];
    for (var i = 0; i < m; i++) {
        for (var j = 0; j < n; j++) { out[i][j] <== a[i][j] * b[i][j]; }
    }
}
component main = matElemMul(2, 2);
"""

ARRAY_ASSIGNMENT = """pragma circom 2.1.0;
template componentA() {
    signal input in[2][2];
    signal output out;
    out <== in[0][0] + in[0][1] + in[1][0] + in[1][1];
}
template componentB() {
    signal input a_in[2][2];
    signal output out;
    component a = componentA();
    a.in <== a_in;
    out <== a.out;
}
component main = componentB();
"""

UNDER_CONSTRAINED = """pragma circom 2.1.0;
template underConstrained() { signal output x; }
component main = underConstrained();
"""

_INFIX = [("mul_2_3", "x2 * x3"), ("idiv_4_3", "x4 \\\\ x3"), ("add_3_4", "x3 + x4"), ("sub_4_1", "x4 - x1"), ("pow_2_4", "x2 ** x4"),
          ("mod_5_3", "x5 % x3"), ("shl_5_1", "x5 << x1"), ("shr_5_1", "x5 >> x1"), ("leq_2_3", "x2 <= x3"), ("leq_3_3", "x3 <= x3"),
          ("leq_4_3", "x4 <= x3"), ("geq_2_3", "x2 >= x3"), ("geq_3_3", "x3 >= x3"), ("geq_4_3", "x4 >= x3"), ("lt_2_3", "x2 < x3"),
          ("lt_3_3", "x3 < x3"), ("lt_4_3", "x4 < x3"), ("gt_2_3", "x2 > x3"), ("gt_3_3", "x3 > x3"), ("gt_4_3", "x4 > x3"),
          ("eq_2_3", "x2 == x3"), ("eq_3_3", "x3 == x3"), ("neq_2_3", "x2 != x3"), ("neq_3_3", "x3 != x3"), ("or_0_1", "x0 || x1"),
          ("and_0_1", "x0 && x1"), ("bit_or_1_3", "x1 | x3"), ("bit_and_1_3", "x1 & x3"), ("bit_xor_1_3", "x1 ^ x3")]
INFIX_OPS = ("pragma circom 2.1.0;\ntemplate infixOps() {\n" + "".join(f"    signal input x{i};\n" for i in range(6))
             + "".join(f"    signal output {n};\n" for n, _ in _INFIX) + "".join(f"    {n} <== {e};\n" for n, e in _INFIX)
             + "}\ncomponent main = infixOps();\n").replace("\\\\", "\\")

_PREFIX = [("negateA", "-a"), ("notA", "!a"), ("notB", "!b"), ("notC", "!c"), ("complementA", "~a"), ("complementB", "~b"), ("complementC", "~c")]
PREFIX_OPS = ("pragma circom 2.1.0;\ntemplate prefixOps() {\n    signal input a; signal input b; signal input c;\n"
              + "".join(f"    signal output {n};\n" for n, _ in _PREFIX) + "".join(f"    {n} <== {e};\n" for n, e in _PREFIX)
              + "}\ncomponent main = prefixOps();\n")

# the CLI's default input (input/circuit.circom): ArgMax(2) out of keras2circom, exercising component arrays inside loops
ARGMAX = """pragma circom 2.0.0;
template Switcher() {
    signal input sel; signal input L; signal input R;
    signal output outL; signal output outR;
    signal aux;
    aux <== (R - L) * sel;
    outL <== aux + L;
    outR <== -aux + R;
}
template ArgMax(n) {
    signal input in[n];
    signal output out;
    signal gts[n];
    component switchers[n + 1];
    component aswitchers[n + 1];
    signal maxs[n + 1];
    signal amaxs[n + 1];
    maxs[0] <== in[0];
    amaxs[0] <== 0;
    for (var i = 0; i < n; i++) {
        gts[i] <== in[i] > maxs[i];
        switchers[i + 1] = Switcher();
        aswitchers[i + 1] = Switcher();
        switchers[i + 1].sel <== gts[i];
        switchers[i + 1].L <== maxs[i];
        switchers[i + 1].R <== in[i];
        aswitchers[i + 1].sel <== gts[i];
        aswitchers[i + 1].L <== amaxs[i];
        aswitchers[i + 1].R <== i;
        amaxs[i + 1] <== aswitchers[i + 1].outL;
        maxs[i + 1] <== switchers[i + 1].outL;
    }
    out <== amaxs[n];
}
component main = ArgMax(N);
"""
