// Host-side sanitizer harness for the packed-stream converters (developer tool): random event streams - valid kinds and deliberately
// invalid ones - through c2a_pack_events / c2a_pack_events_ex (explicit and implicit-operand forms) and back through c2a_unpack_events,
// with exactly-sized heap buffers so that AddressSanitizer sees any overrun.
//   g++ -O1 -g -std=c++17 -pthread -fsanitize=address,undefined -fno-omit-frame-pointer -o /tmp/pack_sanitize tools/pack_sanitize.cpp \
//       tools/front_sanitize_stubs.cpp circom-2-arithc_b200/csrc/c2a_host.cpp && /tmp/pack_sanitize
#include "../include/c2a.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
int main() {
  std::mt19937 rng(12345);
  unsigned long long checked = 0, rejected = 0;
  for (int iter = 0; iter < 20000; ++iter) {
    const unsigned n = rng() % 200;
    std::vector<c2a_event> ev(n);
    unsigned declared = 0;
    const bool dense = rng() % 4 != 0, poison = rng() % 10 == 0;
    for (unsigned i = 0; i < n; ++i) {
      unsigned kind = declared < 2 ? 0 : rng() % 4;
      c2a_event e;
      memset(&e, 0, sizeof e);
      if (kind <= 1) { e.kind = kind; e.a = dense ? declared : (declared * 7 + 3); if (kind == 1) e.b = rng(); ++declared; }
      else if (kind == 2) { e.kind = 2 | ((rng() % 20) << 8); e.a = rng() % declared; e.b = rng() % declared; e.c = (rng() % 3) ? declared - 1 : rng() % declared; }
      else { e.kind = 3; e.a = (rng() % 2) ? declared - 1 : rng() % declared; e.b = rng() % declared; }
      if (!dense && kind >= 2) { e.a = e.a * 7 + 3; e.b = e.b * 7 + 3; if (kind == 2) e.c = e.c * 7 + 3; }
      if (poison && rng() % 50 == 0) e.kind = (rng() % 2) ? (7u | (rng() << 8)) : (2u | (200u << 8));  // invalid kind / gate type
      ev[i] = e;
    }
    for (int allow = 0; allow < 4; ++allow) {
      // sizes first (null outputs), then exactly-sized heap buffers
      c2a_packed_events pk;
      memset(&pk, 0, sizeof pk);
      uint32_t flags = 0, flags2 = 0;
      const unsigned long long nw = c2a_pack_events_ex(ev.data(), n, (uint32_t)allow, nullptr, nullptr, &flags);
      uint32_t* exact_w = (uint32_t*)malloc(nw ? 4 * nw : 4);
      uint8_t* exact_k = (uint8_t*)malloc(n ? n : 1);
      const unsigned long long nw2 = c2a_pack_events_ex(ev.data(), n, (uint32_t)allow, exact_k, exact_w, &flags2);
      if (nw2 != nw || flags2 != flags) { printf("SIZE MISMATCH iter %d allow %d\n", iter, allow); return 1; }
      pk.kinds = exact_k; pk.words = exact_w; pk.n_events = n; pk.n_words = nw; pk.flags = flags;
      c2a_event* back = (c2a_event*)malloc(n ? sizeof(c2a_event) * n : 1);
      int us = c2a_unpack_events(&pk, back);
      if (us != C2A_OK) ++rejected;
      if (us == C2A_OK) {
        for (unsigned i = 0; i < n; ++i) {
          const unsigned k = ev[i].kind & 0xFF;
          bool same = (back[i].kind & 0xFF) == k;
          if (k <= 1) same = same && (!(flags & C2A_PACKED_DENSE_IDS) ? back[i].a == ev[i].a : true);
          else if (k == 2) same = same && back[i].kind == ev[i].kind && back[i].a == ev[i].a && back[i].b == ev[i].b && back[i].c == ev[i].c;
          else if (k == 3) same = same && back[i].a == ev[i].a && back[i].b == ev[i].b;
          if (!same && k <= 3 && !poison) { printf("MISMATCH iter %d allow %d event %u\n", iter, allow, i); return 1; }
        }
        ++checked;
      }
      // a truncated word array must be refused, not read past
      if (nw) { pk.n_words = nw - 1; c2a_unpack_events(&pk, back); }
      free(back); free(exact_w); free(exact_k);
    }
  }
  printf("pack/unpack: %llu round trips checked, %llu packed streams refused by the unpacker (poisoned kinds): no sanitizer report\n", checked, rejected);
  return 0;
}
