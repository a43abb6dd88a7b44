// c2a_internal.h — shared between the device translation units; not part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/c2a.h"

struct c2a_handle {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t stream3 = nullptr;  // copy stream: the payload words of a big packed stream, chunked, while the scatter already consumes them
  cudaEvent_t ev_copy = nullptr;
  cudaStream_t stream2 = nullptr;  // side stream: initialisation of arrays that are only needed later runs next to the kernels before them
  cudaEvent_t ev_side = nullptr;
  cudaEvent_t ev_counts = nullptr;  // side stream: the emitter's early totals have reached the host
  cudaEvent_t ev_side2 = nullptr;  // an earlier point of the side stream (the emitter's outmark clear)
  cudaEvent_t ev_main = nullptr;   // orders the side stream after what the main stream already holds
  // one growable device slab carved per call by a bump allocator (no per-call cudaMalloc)
  char* slab = nullptr;
  size_t slab_bytes = 0;
  size_t slab_used = 0;
  // pinned host staging for scalars and small metadata (I/O pairs)
  uint32_t* h_emit_status = nullptr;  // pinned, 64 u32: the emit's final scalars when its status read is deferred behind the build
  std::string phase_prefix;           // prepended to the phase names recorded while it is set ("emit:" inside c2a_compile_packed*)
  uint32_t* h_pinned = nullptr;
  size_t h_pinned_bytes = 0;
  std::string err;
  uint64_t launches = 0;
  uint64_t relax_fallback_rounds = 0;  // host-synchronised relax rounds beyond the speculative ones (diagnostics)
  // phase timing: CUDA events recorded on `stream`
  struct Phase {
    std::string name;
    cudaEvent_t a, b;
    bool open;
  };
  std::vector<Phase> phases;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_next = 0;
  std::vector<std::pair<std::string, double>> last_ms;
  int nvtx_open = 0;  // NVTX ranges pushed by phase_begin and not yet popped
  bool timing = true;
  std::string timing_only;  // non-empty: events are recorded around this phase only
  // ---- device emitter (c2a_emit.cuh): event staging buffer (separate from the slab) and the resident result
  char* ev_buf = nullptr;
  size_t ev_bytes = 0;
  size_t slab_keep = 0;  // bytes at the start of the slab that slab_reset_keep() preserves (the emitted circuit)
  // compressed streams (c2a_emit_compressed_device): the expanded packed stream on the device, pinned staging for the literal
  // ranges and the chunk tables
  char* cx_buf = nullptr;
  size_t cx_bytes = 0;
  char* cx_pinned = nullptr;
  size_t cx_pinned_bytes = 0;
  struct Emitted {
    bool valid = false;
    bool nos_valid = false;      // node_of_signal[] resident (false for sparse signal ids on the host path)
    size_t gates_off = 0;        // uint4[G], node ids, emission order
    size_t nos_off = 0;          // u32[signal_bound]
    bool prod1_valid = false;    // the producer map (K1) was filled by the emitter's gate kernel: builds skip its memset and k_producer
    size_t prod1_off = 0;        // u32[>= node_count + 1], kept with the circuit (a pure function of the gates)
    uint64_t G = 0;
    uint32_t node_count = 0;
    uint32_t signal_bound = 0;
    const uint32_t* wire = nullptr;  // wire map of the last successful build on this circuit (device; slab scratch or the caller's array)
    bool identity = true;            // that build's DFS order was 0..G-1
  } emitted;
  // single-kernel path (c2a_fused.cuh): double-buffered control block (scalars, grid barrier, look-back slots)
  char* io_buf = nullptr;    // c2a_compile_packed*: the I/O signal lists, uploaded on the side stream ahead of the emit (grow-only)
  size_t io_bytes = 0;
  char* plan_buf = nullptr;  // c2a_plan_shards_device scratch (grow-only, outside the slab)
  size_t plan_bytes = 0;
  char* fused_ctl = nullptr;
  int fused_parity = 0;
  struct c2a_compiler* host_comp = nullptr;  // kept alive when the exact host emitter had to run (sparse ids)
};

namespace c2a {

int fail(c2a_handle* h, int status, const char* fmt, ...);
bool cuda_ok(c2a_handle* h, cudaError_t e, const char* what);

// slab allocator
void slab_reset(c2a_handle* h);       // drops everything, including a resident emitted circuit
void slab_reset_keep(c2a_handle* h);  // keeps the first slab_keep bytes
bool slab_reserve(c2a_handle* h, size_t bytes);  // ensure capacity; reallocation preserves the first slab_keep bytes
void* slab_alloc(c2a_handle* h, size_t bytes);   // 256-byte aligned; nullptr when exhausted
inline size_t align256(size_t b) { return (b + 255) & ~size_t(255); }

// phase timing
void phase_begin(c2a_handle* h, const char* name);
void phase_end(c2a_handle* h);
void phases_collect(c2a_handle* h);  // after a stream sync: fills last_ms, adds "total"
void phases_clear(c2a_handle* h);

// Dependency pairs already on the device -> exact reference DFS order (K5a-c); see c2a_device.cu for the protocol.
struct SortScratch {
  uint32_t* r;
  uint32_t* size_off;  // n+1
  uint8_t* state;
  uint32_t* inq;       // bitmask, (n+31)/32 words
  uint32_t* q0;
  uint32_t* q1;
  uint32_t* heavy;     // n
  uint32_t* tree[6];   // n each: k_tree_blocks (parents, ancestors x2, values x2, subtree sizes)
  unsigned long long* tile_state;   // look-back states of the block-offset scan ...
  unsigned long long* tile_state2;  // ... and of the wire-numbering bitmap scan (one allocation, zeroed together)
  uint32_t* bitmap;                 // first-appearance bitmap over the 3n (position, slot) pairs
  uint32_t* bitmap_pre;             // number of set bits before each bitmap word
  uint32_t bitmap_words;
  size_t tile_state_bytes;
  uint32_t* scalars;   // S_COUNT u32 on the device
};
size_t sort_scratch_bytes(uint64_t n);
bool sort_scratch_carve(c2a_handle* h, uint64_t n, SortScratch* s);
void sort_scalars_reset(c2a_handle* h, const SortScratch& s);
void sort_enqueue_relax(c2a_handle* h, const uint2* d_dep, uint32_t n, const SortScratch& s);
void sort_enqueue_emit(c2a_handle* h, const uint2* d_dep, uint32_t n, const SortScratch& s, uint32_t* d_order);
int sort_status(c2a_handle* h, const uint32_t* host_scalars, uint64_t* err_index);

int grid_for(c2a_handle* h, const void* kernel, int block, uint64_t n);

}  // namespace c2a
