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
# Programs that stress the walker's runtime model (scopes, re-declared variables, the carried return variable, frames with
# more items than the linear scan handles, `const_signal_<v>` spelled by the user, nested component arrays, repeated instances
# of one (template, arguments) pair - the instance memo -, every error class).  The expected streams were produced by the
# string-keyed, memo-free walker this one replaced (a differential run over these programs, the reference's
# tests/circuits/**/*.circom and every fixture: identical events, names, I/O tags and error texts).
WALKER_STRESS = [
    ('pragma circom 2.0.0;\nfunction f(a, b) { var s = 0; for (var i = 0; i < a; i++) { if (i % 2 == 0) { s += i * b; } else { var t = i; s = s + t; } } return s; }\nfunction g(x) { if (x > 3) { return x * 2; } else { return x + 1; } }\ntemplate T(n) { signal input a[n]; signal output o; var acc = f(n, 3); signal p[n]; p[0] <== a[0] * acc; for (var i = 1; i < n; i++) { p[i] <== p[i-1] + a[i] * g(i); } o <== p[n-1]; }\ncomponent main = T(7);',
     0, '', 56, 35, '2989847c867c1ee8'),
    ('pragma circom 2.0.0;\ntemplate Many() { signal input x; signal output y;\n var v0 = 1; var v1 = 2; var v2 = 3; var v3 = 4; var v4 = 5; var v5 = 6; var v6 = 7; var v7 = 8; var v8 = 9; var v9 = 10;\n var w0 = 1; var w1 = 2; var w2 = 3; var w3 = 4; var w4 = 5; var w5 = 6; var w6 = 7; var w7 = 8; var w8 = 9; var w9 = 10;\n signal s0; signal s1; signal s2; signal s3; signal s4; signal s5;\n s0 <== x * v9; s1 <== s0 + w9; s2 <== s1 * 17; s3 <== s2 - v3;\n for (var i = 0; i < 5; i++) { var q0 = i; var q1 = i + 1; var q2 = i + 2; var q3 = q0 * q1; v0 = v0 + q3 + q2; if (i > 2) { var z = 9; w0 += z; } }\n s4 <== s3 * v0; s5 <== s4 + w0; y <== s5 + 12345; }\ncomponent main = Many();',
     0, '', 35, 21, 'dc4a9d61bf187381'),
    ('pragma circom 2.0.0;\ntemplate C() { signal input const_signal_5; signal output o; signal t; t <== const_signal_5 + 5; o <== t * 5; }\ncomponent main = C();',
     0, '', 9, 5, 'fc8220c5601f265d'),
    ('pragma circom 2.0.0;\ntemplate C() { signal input a; signal output o; var const_signal_7 = 3; o <== a * 7; }\ncomponent main = C();',
     102, 'Runtime error: Item already declared', 2, 2, '1335ddd0a8680f1d'),
    ('pragma circom 2.0.0;\ntemplate Inner(k) { signal input in[2][3]; signal output out[2][3]; for (var i = 0; i < 2; i++) { for (var j = 0; j < 3; j++) { out[i][j] <== in[i][j] * k + (i ^ j); } } }\ntemplate Outer() { signal input x[2][3]; signal output y[2][3]; component c[2][2]; for (var a = 0; a < 2; a++) { for (var b = 0; b < 2; b++) { c[a][b] = Inner(a + 2 * b + 1); } }\n c[0][0].in <== x; c[0][1].in <== c[0][0].out; c[1][0].in <== c[0][1].out; c[1][1].in <== c[1][0].out; y <== c[1][1].out; }\ncomponent main = Outer();',
     0, '', 254, 152, 'dac83d1554965219'),
    ('pragma circom 2.0.0;\ntemplate A() { signal input a; signal output b; b <== -a + (~a) + (!a); }\ntemplate B() { signal input in; signal output out; component x = A(); x.a <== in; component y = A(); y.a <== x.b; out <== y.b; }\ncomponent main = B();',
     0, '', 35, 20, '6e86365f3abde6d4'),
    ('pragma circom 2.0.0;\ntemplate A() { signal input a; signal output b; b <== a / 0; }\ncomponent main = A();',
     0, '', 6, 4, '3c3a8312e6dae60f'),
    ('pragma circom 2.0.0;\nfunction bad(x) { return x / 0; }\ntemplate A() { signal input a; signal output b; var q = bad(3); b <== a + q; }\ncomponent main = A();',
     109, 'Operation error: Division by zero', 2, 2, 'c36683aabbe646bf'),
    ('pragma circom 2.0.0;\ntemplate A() { signal input a; signal output b; var e; b <== a + e; }\ncomponent main = A();',
     104, 'Empty data item', 2, 2, 'c36683aabbe646bf'),
    ('pragma circom 2.0.0;\ntemplate A() { signal input a; signal output b; signal a; b <== a; }\ncomponent main = A();',
     102, 'Runtime error: Item already declared', 2, 2, 'c36683aabbe646bf'),
    ('pragma circom 2.0.0;\ntemplate A() { signal input a; signal output b; b <== a + undefined_thing; }\ncomponent main = A();',
     102, 'Runtime error: Item not declared: get_item_data_type: undefined_thing', 2, 2, 'c36683aabbe646bf'),
    ('pragma circom 2.0.0;\ntemplate A() { signal input a; signal output b; component c = Nope(); b <== a; }\ncomponent main = A();',
     112, 'Undefined function or template', 2, 2, 'c36683aabbe646bf'),
    ('pragma circom 2.0.0;\ntemplate A() { signal input a[3]; signal output b; b <== a[1].x; }\ncomponent main = A();',
     102, 'Runtime error: Access Error', 4, 4, 'ba4766d530bac285'),
    ('pragma circom 2.0.0;\ntemplate A() { signal input a[3]; signal output b[2]; b <== a; }\ncomponent main = A();',
     107, 'Invalid data type', 5, 5, '5a9fccab49fdd5c6'),
    ('pragma circom 2.0.0;\ntemplate Sub() { signal input i; signal output o; o <== i * 2; }\ntemplate A() { signal input a; signal output b; component s = Sub(); s.nothere <== a; b <== s.o; }\ncomponent main = A();',
     102, 'Runtime error: Item not declared: get_signal_id: nothere', 8, 6, '1ff8b78ef7172772'),
    ('pragma circom 2.0.0;\ntemplate A() { signal input a; signal output b; b <== a + 4294967296; }\ncomponent main = A();',
     101, 'Parsing error', 2, 2, 'c36683aabbe646bf'),
    ('pragma circom 2.0.0;\nfunction fact(n) { var r = 1; while (n > 1) { r = r * n; n = n - 1; } return r; }\nfunction pw(b, e) { return b ** e; }\ntemplate A(N) { signal input a; signal output b[N]; var c = fact(5) + pw(2, 10) + (7 \\ 2) + (7 % 3) + (1 << 4) + (256 >> 2) + (5 & 3) + (5 | 3) + (5 ^ 3);\n for (var i = 0; i < N; i++) { b[i] <== a * (c + i); } assert(c > 0); }\ncomponent main = A(4);',
     0, '', 21, 13, '1aeb1f5c5b6c1f22'),
    ('pragma circom 2.0.0;\ntemplate A() { signal input a; signal output b; var x = 0; for (var i = 0; i < 3; i++) { var x = i * 10; } b <== a + x; }\ncomponent main = A();',
     0, '', 6, 4, '6dc063b08774486e'),
    ('pragma circom 2.0.0;\ntemplate Leaf() { signal input c; signal output d; d <== c + 1; }\ntemplate Mid(n) { signal input c; signal output d; component l[n]; for (var i = 0; i < n; i++) { l[i] = Leaf(); if (i == 0) { l[i].c <== c; } else { l[i].c <== l[i-1].d; } } d <== l[n-1].d; }\ntemplate Top() { signal input c; signal input cc; signal output d; component m[3]; for (var i = 0; i < 3; i++) { m[i] = Mid(i + 2); m[i].c <== c; } d <== m[0].d + m[1].d + m[2].d + cc + 0; }\ncomponent main = Top();',
     0, '', 88, 50, '82cf7bdb3fe3266b'),
    ('pragma circom 2.0.0;\ntemplate A() { signal input a; signal output b; var i = 0; while (i < 40) { i++; } var big[3][2]; big[2][1] = i; big[0][0] = big[2][1] + 1; b <== a * big[0][0]; var u = big[1][1]; b <== a + u; }\ncomponent main = A();',
     104, 'Empty data item', 6, 4, '914e86bb7f526dda'),
    ('pragma circom 2.0.0;\ntemplate Inner(k) { signal input in[2][3]; signal output out[2][3]; signal mid[2]; for (var i = 0; i < 2; i++) { mid[i] <== in[i][0] * in[i][1]; for (var j = 0; j < 3; j++) { out[i][j] <== in[i][j] * k + mid[i] + 7; } } }\ntemplate Outer() { signal input x[2][3]; signal output y[2][3]; component c[2][2]; for (var a = 0; a < 2; a++) { for (var b = 0; b < 2; b++) { c[a][b] = Inner(3); } }\n c[0][0].in <== x; c[0][1].in <== c[0][0].out; c[1][0].in <== c[0][1].out; c[1][1].in <== c[1][0].out; y <== c[1][1].out; }\ncomponent main = Outer();',
     0, '', 338, 196, '0109a201e98b4095'),
    ('pragma circom 2.0.0;\nfunction fib(n) { var a = 0; var b = 1; for (var i = 0; i < n; i++) { var t = a + b; a = b; b = t; } return a; }\nfunction tri(n) { var s = 0; for (var i = 0; i <= n; i++) { s += fib(i % 7); } return s; }\ntemplate Leaf(a) { signal input c; signal output d; signal e; e <== c * tri(a); d <== e + fib(a) + 0; }\ntemplate Mid(n) { signal input c; signal output d; component l[n]; for (var i = 0; i < n; i++) { l[i] = Leaf(i % 3); if (i == 0) { l[i].c <== c; } else { l[i].c <== l[i-1].d; } } d <== l[n-1].d; }\ntemplate Top() { signal input c; signal output d[6]; component m[6]; for (var i = 0; i < 6; i++) { m[i] = Mid(4 + (i % 2)); m[i].c <== c; d[i] <== m[i].d * 2; } }\ncomponent main = Top();',
     0, '', 427, 241, '81a40adf9ad060f6'),
    ('pragma circom 2.0.0;\ntemplate Z() { signal output o; o <== 5 + 0; }\ntemplate Y() { signal input i; signal output o; component z1 = Z(); component z2 = Z(); component z3 = Z(); o <== i + z1.o + z2.o + z3.o; }\ntemplate X() { signal input i; signal output o; component y[5]; for (var k = 0; k < 5; k++) { y[k] = Y(); if (k == 0) { y[k].i <== i; } else { y[k].i <== y[k-1].o; } } o <== y[4].o; }\ncomponent main = X();',
     0, '', 98, 57, '20061b8c61578208'),
    ('pragma circom 2.0.0;\nfunction f(n) { var r = 0; if (n > 0) { r = f(n - 1); } return r + 1; }\ntemplate A(k) { signal input a; signal output b; var q = f(40) + f(40) + f(k); b <== a * q; }\ntemplate B() { signal input a; signal output b; component x = A(7); component y = A(7); component z = A(41); x.a <== a; y.a <== x.b; z.a <== y.b; b <== z.b; }\ncomponent main = B();',
     0, '', 24, 14, '42e3ac3aa92ca296'),
]
