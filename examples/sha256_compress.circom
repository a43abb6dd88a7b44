pragma circom 2.0.0;
function K(i) {
    var r = 0;
    if (i == 0) { r = 0x428a2f98; }
    if (i == 1) { r = 0x71374491; }
    if (i == 2) { r = 0xb5c0fbcf; }
    if (i == 3) { r = 0xe9b5dba5; }
    if (i == 4) { r = 0x3956c25b; }
    if (i == 5) { r = 0x59f111f1; }
    if (i == 6) { r = 0x923f82a4; }
    if (i == 7) { r = 0xab1c5ed5; }
    if (i == 8) { r = 0xd807aa98; }
    if (i == 9) { r = 0x12835b01; }
    if (i == 10) { r = 0x243185be; }
    if (i == 11) { r = 0x550c7dc3; }
    if (i == 12) { r = 0x72be5d74; }
    if (i == 13) { r = 0x80deb1fe; }
    if (i == 14) { r = 0x9bdc06a7; }
    if (i == 15) { r = 0xc19bf174; }
    if (i == 16) { r = 0xe49b69c1; }
    if (i == 17) { r = 0xefbe4786; }
    if (i == 18) { r = 0x0fc19dc6; }
    if (i == 19) { r = 0x240ca1cc; }
    if (i == 20) { r = 0x2de92c6f; }
    if (i == 21) { r = 0x4a7484aa; }
    if (i == 22) { r = 0x5cb0a9dc; }
    if (i == 23) { r = 0x76f988da; }
    if (i == 24) { r = 0x983e5152; }
    if (i == 25) { r = 0xa831c66d; }
    if (i == 26) { r = 0xb00327c8; }
    if (i == 27) { r = 0xbf597fc7; }
    if (i == 28) { r = 0xc6e00bf3; }
    if (i == 29) { r = 0xd5a79147; }
    if (i == 30) { r = 0x06ca6351; }
    if (i == 31) { r = 0x14292967; }
    if (i == 32) { r = 0x27b70a85; }
    if (i == 33) { r = 0x2e1b2138; }
    if (i == 34) { r = 0x4d2c6dfc; }
    if (i == 35) { r = 0x53380d13; }
    if (i == 36) { r = 0x650a7354; }
    if (i == 37) { r = 0x766a0abb; }
    if (i == 38) { r = 0x81c2c92e; }
    if (i == 39) { r = 0x92722c85; }
    if (i == 40) { r = 0xa2bfe8a1; }
    if (i == 41) { r = 0xa81a664b; }
    if (i == 42) { r = 0xc24b8b70; }
    if (i == 43) { r = 0xc76c51a3; }
    if (i == 44) { r = 0xd192e819; }
    if (i == 45) { r = 0xd6990624; }
    if (i == 46) { r = 0xf40e3585; }
    if (i == 47) { r = 0x106aa070; }
    if (i == 48) { r = 0x19a4c116; }
    if (i == 49) { r = 0x1e376c08; }
    if (i == 50) { r = 0x2748774c; }
    if (i == 51) { r = 0x34b0bcb5; }
    if (i == 52) { r = 0x391c0cb3; }
    if (i == 53) { r = 0x4ed8aa4a; }
    if (i == 54) { r = 0x5b9cca4f; }
    if (i == 55) { r = 0x682e6ff3; }
    if (i == 56) { r = 0x748f82ee; }
    if (i == 57) { r = 0x78a5636f; }
    if (i == 58) { r = 0x84c87814; }
    if (i == 59) { r = 0x8cc70208; }
    if (i == 60) { r = 0x90befffa; }
    if (i == 61) { r = 0xa4506ceb; }
    if (i == 62) { r = 0xbef9a3f7; }
    if (i == 63) { r = 0xc67178f2; }
    return r;
}
template Rotr(n) { signal input x; signal output y; y <== (x >> n) | (x << (32 - n)); }
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
