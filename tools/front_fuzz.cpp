// Mutation fuzzer for the front end (developer tool): random edits of seed programs must end in a status, never in a crash.
//   g++ -O1 -g -std=c++17 -pthread -fsanitize=address,undefined -fno-omit-frame-pointer -o /tmp/front_fuzz tools/front_fuzz.cpp \
//       tools/front_sanitize_stubs.cpp circom-2-arithc_b200/csrc/c2a_front.cpp circom-2-arithc_b200/csrc/c2a_host.cpp
//   /tmp/front_fuzz 20000 examples/poseidon_t3.circom <more seed files>
#include "../include/c2a.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>
int main(int argc, char** argv) {
  std::vector<std::string> seeds;
  for (int i = 2; i < argc; ++i) { std::ifstream f(argv[i]); std::stringstream ss; ss << f.rdbuf(); seeds.push_back(ss.str()); }
  int iters = atoi(argv[1]);
  std::mt19937 rng(argc > 1 ? 777 + atoi(argv[1]) : 1);
  const char* toks[] = {"(", ")", "[", "]", "{", "}", ";", ",", "<==", "==>", "=", "+", "*", "-", "for", "if", "else", "while", "signal", "input", "output", "var", "component", "template", "function", "return", "0", "1", "4294967296", "x", "main", ".", "++", "===", "?", ":", "\"", "/*", "//", "include", "pragma"};
  int st_hist[256] = {0};
  for (int it = 0; it < iters; ++it) {
    std::string s = seeds[rng() % seeds.size()];
    int nm = 1 + rng() % 4;
    for (int m = 0; m < nm && !s.empty(); ++m) {
      size_t pos = rng() % s.size();
      switch (rng() % 5) {
        case 0: s.erase(pos, 1 + rng() % 8); break;
        case 1: s.insert(pos, toks[rng() % (sizeof(toks) / sizeof(*toks))]); break;
        case 2: s[pos] = (char)(32 + rng() % 95); break;
        case 3: { size_t q = rng() % s.size(); size_t len = 1 + rng() % 20; s.insert(pos, s.substr(q, len)); break; }
        default: s.resize(pos); break;
      }
    }
    c2a_program* p = c2a_program_new();
    int st = c2a_program_compile_source(p, s.c_str(), "/nonexistent", nullptr);
    st_hist[st < 0 ? 255 : (st > 254 ? 254 : st)]++;
    if (st == 0) { c2a_packed_events pk; c2a_program_packed(p, &pk); unsigned long long ns = c2a_program_num_signals(p); for (unsigned long long i = 0; i < ns && i < 50; ++i) c2a_program_signal_name(p, (unsigned)i); }
    c2a_program_free(p);
  }
  for (int i = 0; i < 256; ++i) if (st_hist[i]) printf("status %d: %d\n", i, st_hist[i]);
}
