// the two device-side symbols c2a_host.cpp refers to, for host-only builds of the front end (tools/front_sanitize.cpp)
#include "../include/c2a.h"
extern "C" {
int c2a_build_circuit(c2a_handle*, const c2a_gate*, uint64_t, uint32_t, const uint32_t*, uint32_t, const uint32_t*, uint32_t, uint32_t*, uint32_t*, c2a_gate*,
                      uint32_t*, uint64_t*) { return -1; }
const char* c2a_last_error(const c2a_handle*) { return ""; }
}
