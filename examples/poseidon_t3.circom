pragma circom 2.0.0;
function arc(r, lane, t) { return r * t + lane + 1; }
function mds(i, j) { return ((i + 1) * (j + 2)) % 7 + 1; }
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
        if (r < RF \ 2 || r >= RF \ 2 + RP) { full = 1; }
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
component main = Poseidon(3, 8, 57);
