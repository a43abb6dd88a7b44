pragma circom 2.0.0;
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
component main = Main(48, 91);
