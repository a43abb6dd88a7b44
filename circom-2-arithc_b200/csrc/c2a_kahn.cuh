// c2a_kahn.cuh — K4 level-synchronous Kahn frontier + K8/K9 layer-wise sweeps (included by c2a_device.cu).
// (filled in next; the entry points fail loudly until then)
#pragma once
extern "C" {
int c2a_topo_levels(c2a_handle* h, const c2a_gate*, uint64_t, uint32_t, uint32_t*, uint32_t*, uint32_t, uint32_t*, uint64_t*) { return c2a::fail(h, C2A_ERR_INVALID_ARGUMENT, "c2a_topo_levels: not built yet"); }
int c2a_topo_levels_device(c2a_handle* h, const c2a_gate*, uint64_t, uint32_t, uint32_t*, uint32_t*, uint32_t, uint32_t*, uint64_t*) { return c2a::fail(h, C2A_ERR_INVALID_ARGUMENT, "c2a_topo_levels_device: not built yet"); }
int c2a_sweep_masks(c2a_handle* h, const c2a_gate*, uint64_t, uint32_t, const uint32_t*, const uint32_t*, uint32_t, const uint32_t*, uint32_t, uint8_t*, uint32_t*, uint8_t*, uint64_t*) { return c2a::fail(h, C2A_ERR_INVALID_ARGUMENT, "c2a_sweep_masks: not built yet"); }
}
