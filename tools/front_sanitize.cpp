// Host-side sanitizer harness for the front end (developer tool): every program of a file (separated by "\n=====\n") goes through
// compile -> compressed recording -> all signal names -> packed (records carried out on the host) -> 16-byte records -> constants ->
// names in one call, without and with a host emitter attached (then also the native circuit report).
//   g++ -O1 -g -std=c++17 -pthread -fsanitize=address,undefined -fno-omit-frame-pointer -o /tmp/front_sanitize tools/front_sanitize.cpp \
//       tools/front_sanitize_stubs.cpp circom-2-arithc_b200/csrc/c2a_front.cpp circom-2-arithc_b200/csrc/c2a_host.cpp && /tmp/front_sanitize programs.txt
#include "../include/c2a.h"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
// runs every program (separated by "\n=====\n") through: compile (lazy) -> compressed -> names of all signals -> packed (materialise)
// -> events -> constants; and compile with a host emitter attached
int main(int argc, char** argv) {
  std::ifstream f(argv[1]); std::stringstream ss; ss << f.rdbuf(); std::string all = ss.str();
  std::vector<std::string> cases; size_t at = 0;
  while (true) { size_t k = all.find("\n=====\n", at); cases.push_back(all.substr(at, k == std::string::npos ? k : k - at)); if (k == std::string::npos) break; at = k + 7; }
  unsigned long long total = 0;
  for (auto& src : cases) {
    for (int with = 0; with < 2; ++with) {
      c2a_program* p = c2a_program_new();
      c2a_compiler* comp = with ? c2a_compiler_new() : nullptr;
      int st = c2a_program_compile_source(p, src.c_str(), nullptr, comp);
      c2a_compressed_events cx; c2a_program_compressed(p, &cx);
      unsigned long long ns = c2a_program_num_signals(p);
      for (unsigned long long i = 0; i < ns; ++i) total += strlen(c2a_program_signal_name(p, (unsigned)i));
      c2a_packed_events pk; c2a_program_packed(p, &pk);
      for (unsigned long long i = 0; i < pk.n_events; ++i) total += pk.kinds[i];
      for (unsigned long long i = 0; i < pk.n_words; ++i) total += pk.words[i];
      const c2a_event* ev = c2a_program_events(p);
      for (unsigned long long i = 0; i < c2a_program_num_events(p); ++i) total += ev[i].a + ev[i].b + ev[i].c;
      for (unsigned long long i = 0; i < c2a_program_num_constants(p); ++i) total += c2a_program_constant_signals(p)[i] + c2a_program_constant_values(p)[i];
      std::vector<unsigned> ids(ns); for (unsigned i = 0; i < ns; ++i) ids[i] = i;
      unsigned long long need = c2a_program_signal_names(p, ids.data(), ns, nullptr, 0);
      std::vector<char> buf(need + 1); c2a_program_signal_names(p, ids.data(), ns, buf.data(), need);
      if (comp) { const char* rep = c2a_circuit_report_json(comp, "sint"); total += strlen(rep); c2a_compiler_free(comp); }
      printf("st=%d events=%llu replays=%llu gen=%u\n", st, (unsigned long long)pk.n_events, (unsigned long long)cx.n_replays, cx.max_gen);
      c2a_program_free(p);
    }
  }
  printf("checksum %llu\n", total);
}
