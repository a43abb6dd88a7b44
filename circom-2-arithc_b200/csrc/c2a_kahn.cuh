// c2a_kahn.cuh — K3/K4: consumer CSR + level-synchronous Kahn frontier, and the opt-in layer-wise sweeps
// (K8 constant-fold mask, K9 dead-gate mask).  Included by c2a_device.cu (single translation unit).
//
// The dependency relation is the reference's (src/compiler.rs:401-421): gate g depends on the LAST producer
// of its lh and rh nodes.  Kahn yields a valid topological order and the levels, NOT the reference's DFS
// post-order (that is K5, c2a_device.cu); it is used for cycle screening, the sweeps and evaluators.
//
//   K3  k_kahn_count / k_scan_u32 / k_kahn_fill   consumer CSR  row_off[G+1], col[<=2G]  (multiplicity kept)
//   K4  k_kahn_frontier   ONE persistent cooperative launch for all levels:
//         * the output array level_order[] is the BFS queue itself: [lo,hi) is the current frontier
//         * in-degree decrements are atomicSub on indeg[]; newly ready gates are appended with one
//           warp-aggregated atomicAdd per warp (ballot + popc)
//         * rows of >= kBulkRow consumers are staged into shared memory with a TMA bulk copy
//           (cp.async.bulk + mbarrier) by the whole CTA
//         * one grid barrier per level; when the frontier is small (<= kSmallFrontier) CTA 0 runs consecutive
//           levels alone with __syncthreads() only while the other CTAs wait at the barrier
#pragma once

namespace c2a {

constexpr int kKahnBlock = 512;
constexpr uint32_t kSmallFrontier = 2048;
constexpr uint32_t kBulkRow = 1024;       // consumers; rows at least this long go through the TMA bulk path
constexpr uint32_t kBulkChunk = 4096;     // u32 entries staged per bulk copy (16 KB)
enum { KC_TAIL = 0, KC_LO = 1, KC_HI = 2, KC_LEVEL = 3, KC_ARRIVE = 4, KC_RELEASE = 5, KC_ERRMIN = 6, KC_OVERFLOW = 7, KC_BIGN = 8, KC_COUNT = 16 };

__global__ void __launch_bounds__(kBlock) k_kahn_count(const uint2* __restrict__ dep, uint32_t G, uint32_t* __restrict__ indeg,
                                                       uint32_t* __restrict__ cnt) {
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < G; g += gridDim.x * kBlock) {
    uint2 d = dep[g];
    uint32_t n = 0;
    if (d.x != kNone) { atomicAdd(cnt + d.x, 1u); ++n; }
    if (d.y != kNone) { atomicAdd(cnt + d.y, 1u); ++n; }  // lh == rh: the same consumer twice, decremented twice
    indeg[g] = n;
  }
}

__global__ void __launch_bounds__(kBlock) k_kahn_fill(const uint2* __restrict__ dep, uint32_t G, const uint32_t* __restrict__ row_off,
                                                      uint32_t* __restrict__ cursor, uint32_t* __restrict__ col) {
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < G; g += gridDim.x * kBlock) {
    uint2 d = dep[g];
    if (d.x != kNone) col[row_off[d.x] + atomicAdd(cursor + d.x, 1u)] = g;
    if (d.y != kNone) col[row_off[d.y] + atomicAdd(cursor + d.y, 1u)] = g;
  }
}

// ---- grid barrier with a serial section run by the last CTA to arrive --------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Advance the frontier window: called by exactly one thread when every append of the level is visible.
__device__ __forceinline__ void kahn_advance(uint32_t* ctrl, uint32_t* __restrict__ level_off, uint32_t level_cap) {
  uint32_t hi = ctrl[KC_HI], tail = ld_acquire_u32(ctrl + KC_TAIL), lvl = ctrl[KC_LEVEL] + 1;
  ctrl[KC_LO] = hi;
  ctrl[KC_HI] = tail;
  ctrl[KC_LEVEL] = lvl;
  if (lvl <= level_cap) level_off[lvl] = hi; else ctrl[KC_OVERFLOW] = 1;
}

template <bool kAdvance>
__device__ __forceinline__ void kahn_grid_barrier(uint32_t* ctrl, uint32_t& gen, uint32_t nblocks, uint32_t* level_off, uint32_t level_cap,
                                                  bool advance) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    uint32_t old = atomicAdd(ctrl + KC_ARRIVE, 1u);
    if (old == nblocks * (gen + 1) - 1) {  // last to arrive: everything of this level is visible
      if (kAdvance && advance) kahn_advance(ctrl, level_off, level_cap);
      __threadfence();
      st_release_u32(ctrl + KC_RELEASE, gen + 1);
    } else {
      uint32_t spins = 0;
      while (ld_acquire_u32(ctrl + KC_RELEASE) < gen + 1) {
        __nanosleep(32);
        if (++spins > (1u << 27)) __trap();  // ~10 s: a CTA never arrived (fail loudly, do not hang)
      }
    }
  }
  ++gen;
  __syncthreads();
}

// warp-aggregated append of `item` for lanes with ready != 0 (all 32 lanes must call)
__device__ __forceinline__ void warp_append(bool ready, uint32_t item, uint32_t* __restrict__ queue, uint32_t* __restrict__ tail) {
  uint32_t m = __ballot_sync(0xFFFFFFFFu, ready);
  if (!m) return;
  int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(tail, (uint32_t)__popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
  if (ready) queue[base + __popc(m & ((1u << lane) - 1))] = item;
}

__device__ __forceinline__ void kahn_relax_consumer(bool valid, uint32_t c, uint32_t* __restrict__ indeg, uint32_t* __restrict__ queue,
                                                    uint32_t* __restrict__ tail) {
  bool ready = false;
  if (valid) ready = atomicSub(indeg + c, 1u) == 1u;
  warp_append(ready, c, queue, tail);
}

// TMA 1-D bulk copy global -> shared, completion on an mbarrier (cp.async.bulk; SASS: UBLKCP)
__device__ __forceinline__ void bulk_load_row(uint32_t* smem_dst, const uint32_t* gsrc, uint32_t bytes, unsigned long long* mbar, uint32_t phase) {
  uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
  }
  // all threads wait for the phase to complete
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
  }
}

// Process frontier items [lo,hi) with the threads of this CTA that are given (stride = number of cooperating
// threads across the grid).  Rows shorter than kBulkRow: one lane per item, the warp walks the rows in lock
// step (ballots stay converged).  Longer rows: recorded in s_big and handled by the whole CTA through smem.
__device__ __forceinline__ void kahn_process(uint32_t lo, uint32_t hi, uint32_t first, uint32_t stride, const uint32_t* __restrict__ row_off,
                                             const uint32_t* __restrict__ col, uint32_t* __restrict__ indeg, uint32_t* __restrict__ queue,
                                             uint32_t* ctrl, uint32_t* s_big, uint32_t* s_bign, uint32_t* s_stage, unsigned long long* s_mbar,
                                             uint32_t& mbar_phase) {
  const int lane = threadIdx.x & 31;
  // every warp iterates the same number of times over its slice so that the ballots are full-warp
  uint32_t n = hi - lo;
  uint32_t iters = (n + stride - 1) / stride;
  for (uint32_t it = 0; it < iters; ++it) {
    uint32_t i = lo + it * stride + first;
    uint32_t beg = 0, len = 0;
    if (i < hi) {
      uint32_t g = queue[i];
      beg = row_off[g];
      len = row_off[g + 1] - beg;
      if (len >= kBulkRow) {  // defer to the CTA-wide bulk path
        uint32_t slot = atomicAdd(s_bign, 1u);
        if (slot < 64) { s_big[2 * slot] = beg; s_big[2 * slot + 1] = len; len = 0; }
        // more than 64 long rows in one CTA pass: fall through and walk it here (correct, just slower)
      }
    }
    // short rows: lock-step walk
    uint32_t maxlen = len;
#pragma unroll
    for (int o = 16; o; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xFFFFFFFFu, maxlen, o));
    if (maxlen <= 32) {
      for (uint32_t j = 0; j < maxlen; ++j) {
        bool v = j < len;
        uint32_t c = v ? col[beg + j] : 0u;
        kahn_relax_consumer(v, c, indeg, queue, ctrl + KC_TAIL);
      }
    } else {  // medium rows: the warp takes the lanes' rows one after the other, 32 consumers at a time
      uint32_t has = __ballot_sync(0xFFFFFFFFu, len > 0);
      while (has) {
        int src = __ffs(has) - 1;
        has &= has - 1;
        uint32_t b = __shfl_sync(0xFFFFFFFFu, beg, src), l = __shfl_sync(0xFFFFFFFFu, len, src);
        for (uint32_t j = 0; j < l; j += 32) {
          bool v = j + lane < l;
          uint32_t c = v ? col[b + j + lane] : 0u;
          kahn_relax_consumer(v, c, indeg, queue, ctrl + KC_TAIL);
        }
      }
    }
  }
  __syncthreads();
  // long rows: CTA-wide, staged through shared memory by TMA bulk copies of 16-byte aligned chunks
  uint32_t nb = min(*s_bign, 64u);
  for (uint32_t r = 0; r < nb; ++r) {
    uint32_t beg = s_big[2 * r], len = s_big[2 * r + 1];
    uint32_t pos = beg, end = beg + len;
    // unaligned head straight from global
    uint32_t head_end = min(end, (pos + 3u) & ~3u);
    for (uint32_t j = pos + threadIdx.x; j < ((head_end - pos + 31) / 32) * 32 + pos; j += blockDim.x) {
      bool v = j < head_end;
      kahn_relax_consumer(v, v ? col[j] : 0u, indeg, queue, ctrl + KC_TAIL);
    }
    pos = head_end;
    while (pos + 4 <= end) {
      uint32_t cnt = min((end - pos) & ~3u, kBulkChunk);
      bulk_load_row(s_stage, col + pos, cnt * 4, s_mbar, mbar_phase);
      mbar_phase ^= 1;
      for (uint32_t j = threadIdx.x; j < ((cnt + 31) / 32) * 32; j += blockDim.x) {
        bool v = j < cnt;
        kahn_relax_consumer(v, v ? s_stage[j] : 0u, indeg, queue, ctrl + KC_TAIL);
      }
      __syncthreads();  // everyone is done with the staging buffer
      pos += cnt;
    }
    for (uint32_t j = pos + threadIdx.x; j < ((end - pos + 31) / 32) * 32 + pos; j += blockDim.x) {  // tail (<4 entries)
      bool v = j < end;
      kahn_relax_consumer(v, v ? col[j] : 0u, indeg, queue, ctrl + KC_TAIL);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *s_bign = 0;
  __syncthreads();
}

__global__ void __launch_bounds__(kKahnBlock) k_kahn_frontier(uint32_t G, const uint32_t* __restrict__ row_off, const uint32_t* __restrict__ col,
                                                              uint32_t* __restrict__ indeg, uint32_t* __restrict__ queue,
                                                              uint32_t* __restrict__ level_off, uint32_t level_cap, uint32_t* ctrl) {
  __shared__ uint32_t s_big[128];
  __shared__ uint32_t s_bign;
  __shared__ __align__(16) uint32_t s_stage[kBulkChunk];
  __shared__ __align__(8) unsigned long long s_mbar;
  __shared__ uint32_t s_lo, s_hi;
  const uint32_t nblocks = gridDim.x;
  uint32_t gen = 0, mbar_phase = 0;
  if (threadIdx.x == 0) {
    s_bign = 0;
    uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // level 0: gates without dependencies, in ascending index order per warp chunk
  {
    uint32_t stride = nblocks * kKahnBlock;
    uint32_t iters = (G + stride - 1) / stride;
    for (uint32_t it = 0; it < iters; ++it) {
      uint32_t g = it * stride + blockIdx.x * kKahnBlock + threadIdx.x;
      bool ready = g < G && indeg[g] == 0;
      warp_append(ready, g, queue, ctrl + KC_TAIL);
    }
  }
  // the last arriver publishes [0, tail) as level 0  (ctrl starts as lo=hi=0, level=-1 -> advance gives level 0)
  kahn_grid_barrier<true>(ctrl, gen, nblocks, level_off, level_cap, true);
  while (true) {
    uint32_t lo = ld_acquire_u32(ctrl + KC_LO), hi = ld_acquire_u32(ctrl + KC_HI);
    if (lo == hi) break;
    bool small = hi - lo <= kSmallFrontier;
    if (!small) {
      kahn_process(lo, hi, blockIdx.x * kKahnBlock + threadIdx.x, nblocks * kKahnBlock, row_off, col, indeg, queue, ctrl, s_big, &s_bign, s_stage,
                   &s_mbar, mbar_phase);
    } else if (blockIdx.x == 0) {
      // small-frontier mode: this CTA alone runs consecutive levels; no grid barrier in between
      while (true) {
        kahn_process(lo, hi, threadIdx.x, kKahnBlock, row_off, col, indeg, queue, ctrl, s_big, &s_bign, s_stage, &s_mbar, mbar_phase);
        if (threadIdx.x == 0) {
          __threadfence();
          kahn_advance(ctrl, level_off, level_cap);
          s_lo = ctrl[KC_LO];
          s_hi = ctrl[KC_HI];
        }
        __syncthreads();
        lo = s_lo;
        hi = s_hi;
        __syncthreads();
        if (lo == hi || hi - lo > kSmallFrontier) break;
      }
    }
    kahn_grid_barrier<true>(ctrl, gen, nblocks, level_off, level_cap, !small);
  }
}

// after the frontier ran dry: gates that were never released sit on or behind a cycle
__global__ void __launch_bounds__(kBlock) k_kahn_leftover(const uint32_t* __restrict__ indeg, uint32_t G, uint32_t* ctrl) {
  uint32_t m = kNone;
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < G; g += gridDim.x * kBlock)
    if (indeg[g] != 0) { m = g; break; }
  if (m != kNone) atomicMin(ctrl + KC_ERRMIN, m);
}

// ---------------------------------------------------------------------------------------------------
// K8 / K9: layer-wise sweeps over the Kahn levels (opt-in; never alter build_circuit output)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool exec_op_u32(uint32_t op, uint32_t a, uint32_t b, uint32_t* r) {
  // src/process.rs:649-750 execute_op; returns false where the reference errors (or would overflow-panic)
  switch (op) {
    case C2A_AMul: { unsigned long long v = (unsigned long long)a * b; *r = (uint32_t)v; return !(v >> 32); }
    case C2A_ADiv: case C2A_AIntDiv: if (!b) return false; *r = a / b; return true;
    case C2A_AAdd: { unsigned long long v = (unsigned long long)a + b; *r = (uint32_t)v; return !(v >> 32); }
    case C2A_ASub: if (a < b) return false; *r = a - b; return true;
    case C2A_APow: {
      unsigned long long base = a, acc = 1; uint32_t e = b; bool ovf = false;
      while (e) { if (e & 1) { acc *= base; if (acc >> 32) { ovf = true; acc &= 0xFFFFFFFFull; } } e >>= 1; if (e) { base *= base; if (base >> 32) { ovf = true; base &= 0xFFFFFFFFull; } } }
      *r = (uint32_t)acc; return !ovf;
    }
    case C2A_AMod: if (!b) return false; *r = a % b; return true;
    case C2A_AShiftL: if (b >= 32) return false; *r = a << b; return true;
    case C2A_AShiftR: if (b >= 32) return false; *r = a >> b; return true;
    case C2A_ALEq: *r = a <= b; return true;
    case C2A_AGEq: *r = a >= b; return true;
    case C2A_ALt: *r = a < b; return true;
    case C2A_AGt: *r = a > b; return true;
    case C2A_AEq: *r = a == b; return true;
    case C2A_ANeq: *r = a != b; return true;
    case C2A_ABoolOr: *r = (a != 0 || b != 0); return true;
    case C2A_ABoolAnd: *r = (a != 0 && b != 0); return true;
    case C2A_ABitOr: *r = a | b; return true;
    case C2A_ABitAnd: *r = a & b; return true;
    case C2A_AXor: *r = a ^ b; return true;
  }
  return false;
}

// node_const[node]: bit0 = value known; node_val[node] = value.  One launch per level (forward).
__global__ void __launch_bounds__(kBlock) k_fold_level(const uint4* __restrict__ gates, const uint32_t* __restrict__ level_order, uint32_t lo,
                                                       uint32_t hi, const uint32_t* __restrict__ prod1, uint8_t* __restrict__ node_const,
                                                       uint32_t* __restrict__ node_val, uint8_t* __restrict__ const_mask,
                                                       uint32_t* __restrict__ const_value) {
  for (uint32_t i = lo + blockIdx.x * kBlock + threadIdx.x; i < hi; i += gridDim.x * kBlock) {
    uint32_t g = level_order[i];
    uint4 gt = gates[g];
    uint32_t r = 0;
    bool c = node_const[gt.y] && node_const[gt.z] && exec_op_u32(gt.x, node_val[gt.y], node_val[gt.z], &r);
    const_mask[g] = c;
    const_value[g] = c ? r : 0u;
    if (prod1[gt.w] == g + 1) {  // this gate is the node's producer (last writer wins, compiler.rs:403-406)
      node_const[gt.w] = c;
      node_val[gt.w] = r;
    }
  }
}

__global__ void __launch_bounds__(kBlock) k_set_node_consts(const uint32_t* __restrict__ nodes, const uint32_t* __restrict__ vals, uint32_t n,
                                                            uint32_t node_bound, uint8_t* __restrict__ node_const, uint32_t* __restrict__ node_val) {
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock)
    if (nodes[i] < node_bound) { node_const[nodes[i]] = 1; node_val[nodes[i]] = vals[i]; }
}

__global__ void __launch_bounds__(kBlock) k_mark_nodes(const uint32_t* __restrict__ nodes, uint32_t n, uint32_t node_bound, uint8_t* __restrict__ mark) {
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock)
    if (nodes[i] < node_bound) mark[nodes[i]] = 1;
}

// reverse sweep: a gate is live when it produces an output node or feeds a live gate
__global__ void __launch_bounds__(kBlock) k_live_level(const uint4* __restrict__ gates, const uint32_t* __restrict__ level_order, uint32_t lo,
                                                       uint32_t hi, const uint32_t* __restrict__ prod1, const uint32_t* __restrict__ row_off,
                                                       const uint32_t* __restrict__ col, const uint8_t* __restrict__ out_mark,
                                                       uint8_t* __restrict__ live) {
  for (uint32_t i = lo + blockIdx.x * kBlock + threadIdx.x; i < hi; i += gridDim.x * kBlock) {
    uint32_t g = level_order[i];
    uint4 gt = gates[g];
    bool l = out_mark[gt.w] && prod1[gt.w] == g + 1;
    for (uint32_t j = row_off[g]; !l && j < row_off[g + 1]; ++j) l = live[col[j]] != 0;
    live[g] = l;
  }
}

__global__ void __launch_bounds__(kBlock) k_invert_u8(const uint8_t* __restrict__ a, uint8_t* __restrict__ b, uint32_t n) {
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) b[i] = a[i] ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------------
struct KahnBuffers {
  uint32_t* prod1;
  uint2* dep;
  uint32_t* indeg;
  uint32_t* row_off;  // G+1
  uint32_t* cursor;
  uint32_t* col;      // 2G
  unsigned long long* tile_state;
  uint32_t* scalars;
  uint32_t* ctrl;
};

static size_t kahn_scratch_bytes(uint64_t G, uint32_t node_bound) {
  return align256(4 * (size_t)node_bound) + align256(8 * G) + align256(4 * G) + align256(4 * (G + 1)) + align256(4 * G) + align256(8 * G + 4) +
         align256(8 * (size_t)(scan_tiles(G, kScanItems) + 1)) + align256(4 * S_COUNT) + align256(4 * KC_COUNT);
}

static bool kahn_carve(c2a_handle* h, uint64_t G, uint32_t node_bound, KahnBuffers* b) {
  b->prod1 = (uint32_t*)slab_alloc(h, 4 * (size_t)node_bound);
  b->dep = (uint2*)slab_alloc(h, 8 * G);
  b->indeg = (uint32_t*)slab_alloc(h, 4 * G);
  b->row_off = (uint32_t*)slab_alloc(h, 4 * (G + 1));
  b->cursor = (uint32_t*)slab_alloc(h, 4 * G);
  b->col = (uint32_t*)slab_alloc(h, 8 * G + 4);
  b->tile_state = (unsigned long long*)slab_alloc(h, 8 * (size_t)(scan_tiles(G, kScanItems) + 1));
  b->scalars = (uint32_t*)slab_alloc(h, 4 * S_COUNT);
  b->ctrl = (uint32_t*)slab_alloc(h, 4 * KC_COUNT);
  return b->ctrl != nullptr;
}

// Runs K1,K2,K3,K4 on device-resident gates.  d_level_order[G], d_level_off[level_cap+1] on the device.
static int kahn_core(c2a_handle* h, const uint4* d_gates, uint32_t G, uint32_t node_bound, const KahnBuffers& b, uint32_t* d_level_order,
                     uint32_t* d_level_off, uint32_t level_cap, uint32_t* n_levels, uint64_t* err_index) {
  cudaStream_t st = h->stream;
  uint32_t* hp = h->h_pinned;
  for (int i = 0; i < S_COUNT; ++i) hp[i] = 0;
  for (int i = 0; i < KC_COUNT; ++i) hp[S_COUNT + i] = 0;
  hp[S_COUNT + KC_LEVEL] = 0xFFFFFFFFu;  // first advance -> level 0
  hp[S_COUNT + KC_ERRMIN] = kNone;
  cudaMemcpyAsync(b.scalars, hp, 4 * S_COUNT, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(b.ctrl, hp + S_COUNT, 4 * KC_COUNT, cudaMemcpyHostToDevice, st);
  phase_begin(h, "init");
  cudaMemsetAsync(b.prod1, 0, 4 * (size_t)node_bound, st);
  cudaMemsetAsync(b.row_off, 0, 4 * ((size_t)G + 1), st);
  cudaMemsetAsync(b.cursor, 0, 4 * (size_t)G, st);
  phase_end(h);
  if (G) {
    phase_begin(h, "k_producer");
    LAUNCH(h, k_producer, grid_for(h, (const void*)k_producer, kBlock, G), kBlock, d_gates, G, node_bound, b.prod1, b.scalars);
    phase_end(h);
    phase_begin(h, "k_deps");
    LAUNCH(h, k_deps, grid_for(h, (const void*)k_deps, kBlock, G), kBlock, d_gates, G, node_bound, b.prod1, b.dep, b.scalars);
    phase_end(h);
    phase_begin(h, "k_kahn_count");
    LAUNCH(h, k_kahn_count, grid_for(h, (const void*)k_kahn_count, kBlock, G), kBlock, b.dep, G, b.indeg, b.row_off);
    phase_end(h);
    uint32_t tiles = scan_tiles(G, kScanItems);
    cudaMemsetAsync(b.tile_state, 0, 8 * (size_t)tiles, st);
    phase_begin(h, "k_scan_u32");
    LAUNCH(h, k_scan_u32_t<false>, tiles, kBlock, b.row_off, b.row_off, G, b.tile_state, (uint32_t*)nullptr, (const uint32_t*)nullptr, 0);
    phase_end(h);
    phase_begin(h, "k_kahn_fill");
    LAUNCH(h, k_kahn_fill, grid_for(h, (const void*)k_kahn_fill, kBlock, G), kBlock, b.dep, G, b.row_off, b.cursor, b.col);
    phase_end(h);
    // persistent cooperative launch: every CTA must be resident for the grid barrier
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_kahn_frontier, kKahnBlock, 0) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
    int grid = h->num_sms * std::min(occ, 2);
    const uint32_t* row_off = b.row_off;
    const uint32_t* col = b.col;
    uint32_t* indeg = b.indeg;
    uint32_t* ctrl = b.ctrl;
    uint32_t Gv = G;
    void* args[] = {&Gv, &row_off, &col, &indeg, &d_level_order, &d_level_off, &level_cap, &ctrl};
    phase_begin(h, "k_kahn_frontier");
    if (!cuda_ok(h, cudaLaunchCooperativeKernel((const void*)k_kahn_frontier, dim3(grid), dim3(kKahnBlock), args, 0, st), "kahn cooperative launch")) return C2A_ERR_CUDA;
    h->launches++;
    phase_end(h);
    LAUNCH(h, k_kahn_leftover, grid_for(h, (const void*)k_kahn_leftover, kBlock, G), kBlock, b.indeg, G, b.ctrl);
  }
  cudaMemcpyAsync(hp + S_COUNT, b.ctrl, 4 * KC_COUNT, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(hp, b.scalars, 4 * S_COUNT, cudaMemcpyDeviceToHost, st);
  if (!cuda_ok(h, cudaStreamSynchronize(st), "kahn sync")) return C2A_ERR_CUDA;
  if (!cuda_ok(h, cudaGetLastError(), "kahn kernels")) return C2A_ERR_CUDA;
  if (hp[S_FLAGS] & F_BAD) return fail(h, C2A_ERR_INVALID_ARGUMENT, "a gate references a node id >= node_bound (%u)", node_bound);
  uint32_t tail = hp[S_COUNT + KC_TAIL];
  if (tail != G) {
    if (err_index) *err_index = hp[S_COUNT + KC_ERRMIN];
    return fail(h, C2A_ERR_CYCLIC_DEPENDENCY, "%u of %u gates are on or behind a dependency cycle (smallest index %u)", G - tail, G, hp[S_COUNT + KC_ERRMIN]);
  }
  if (hp[S_COUNT + KC_OVERFLOW]) return fail(h, C2A_ERR_INVALID_ARGUMENT, "more levels than level_cap (%u)", level_cap);
  // the final advance recorded an empty level: levels = KC_LEVEL (0-based index of that empty level)
  uint32_t levels = G ? hp[S_COUNT + KC_LEVEL] : 0;
  if (n_levels) *n_levels = levels;
  return C2A_OK;
}

}  // namespace c2a

using namespace c2a;

extern "C" {

int c2a_topo_levels_device(c2a_handle* h, const c2a_gate* d_gates, uint64_t G, uint32_t node_bound, uint32_t* d_level_order, uint32_t* d_level_off,
                           uint32_t level_cap, uint32_t* n_levels, uint64_t* err_index) {
  int st = check_sizes(h, G, node_bound);
  if (st) return st;
  if ((G && (!d_level_order || !d_gates)) || !d_level_off) return fail(h, C2A_ERR_INVALID_ARGUMENT, "null argument");
  phases_clear(h);
  slab_reset(h);
  if (!slab_reserve(h, kahn_scratch_bytes(G, node_bound))) return C2A_ERR_NO_MEMORY;
  KahnBuffers b;
  if (!kahn_carve(h, G, node_bound, &b)) return fail(h, C2A_ERR_NO_MEMORY, "scratch slab exhausted");
  st = kahn_core(h, (const uint4*)d_gates, (uint32_t)G, node_bound, b, d_level_order, d_level_off, level_cap, n_levels, err_index);
  cudaStreamSynchronize(h->stream);
  phases_collect(h);
  return st;
}

int c2a_topo_levels(c2a_handle* h, const c2a_gate* gates, uint64_t G, uint32_t node_bound, uint32_t* level_order, uint32_t* level_off,
                    uint32_t level_cap, uint32_t* n_levels, uint64_t* err_index) {
  int st = check_sizes(h, G, node_bound);
  if (st) return st;
  if ((G && (!level_order || !gates)) || !level_off) return fail(h, C2A_ERR_INVALID_ARGUMENT, "null argument");
  phases_clear(h);
  slab_reset(h);
  if (!slab_reserve(h, kahn_scratch_bytes(G, node_bound) + align256(16 * G) + align256(4 * G) + align256(4 * ((size_t)level_cap + 2)))) return C2A_ERR_NO_MEMORY;
  KahnBuffers b;
  bool ok = kahn_carve(h, G, node_bound, &b);
  uint4* d_gates = (uint4*)slab_alloc(h, 16 * G);
  uint32_t* d_lo = (uint32_t*)slab_alloc(h, 4 * G);
  uint32_t* d_off = (uint32_t*)slab_alloc(h, 4 * ((size_t)level_cap + 2));
  if (!ok || !d_gates || !d_lo || !d_off) return fail(h, C2A_ERR_NO_MEMORY, "scratch slab exhausted");
  cudaStream_t s = h->stream;
  if (G) cudaMemcpyAsync(d_gates, gates, 16 * G, cudaMemcpyHostToDevice, s);
  cudaMemsetAsync(d_off, 0, 4 * ((size_t)level_cap + 2), s);
  uint32_t nl = 0;
  st = kahn_core(h, d_gates, (uint32_t)G, node_bound, b, d_lo, d_off, level_cap, &nl, err_index);
  if (st == C2A_OK) {
    if (G) cudaMemcpyAsync(level_order, d_lo, 4 * G, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(level_off, d_off, 4 * ((size_t)nl + 1), cudaMemcpyDeviceToHost, s);
    if (n_levels) *n_levels = nl;
  }
  if (!cuda_ok(h, cudaStreamSynchronize(s), "levels D2H") && st == C2A_OK) st = C2A_ERR_CUDA;
  phases_collect(h);
  return st;
}

int c2a_sweep_masks(c2a_handle* h, const c2a_gate* gates, uint64_t G, uint32_t node_bound, const uint32_t* const_nodes, const uint32_t* const_values,
                    uint32_t n_const, const uint32_t* output_nodes, uint32_t n_out, uint8_t* const_mask, uint32_t* const_value, uint8_t* dead_mask,
                    uint64_t* err_index) {
  int st = check_sizes(h, G, node_bound);
  if (st) return st;
  phases_clear(h);
  slab_reset(h);
  size_t need = kahn_scratch_bytes(G, node_bound) + align256(16 * G) + align256(4 * G) + align256(4 * (G + 2)) + 2 * align256(node_bound) +
                align256(4 * (size_t)node_bound) + 3 * align256(G) + align256(4 * G) + 2 * align256(4 * (size_t)n_const + 4) + align256(4 * (size_t)n_out + 4);
  if (!slab_reserve(h, need)) return C2A_ERR_NO_MEMORY;
  KahnBuffers b;
  bool ok = kahn_carve(h, G, node_bound, &b);
  uint4* d_gates = (uint4*)slab_alloc(h, 16 * G);
  uint32_t* d_lo = (uint32_t*)slab_alloc(h, 4 * G);
  uint32_t* d_off = (uint32_t*)slab_alloc(h, 4 * (G + 2));
  uint8_t* node_const = (uint8_t*)slab_alloc(h, node_bound);
  uint8_t* out_mark = (uint8_t*)slab_alloc(h, node_bound);
  uint32_t* node_val = (uint32_t*)slab_alloc(h, 4 * (size_t)node_bound);
  uint8_t* d_cmask = (uint8_t*)slab_alloc(h, G);
  uint8_t* d_live = (uint8_t*)slab_alloc(h, G);
  uint8_t* d_dead = (uint8_t*)slab_alloc(h, G);
  uint32_t* d_cval = (uint32_t*)slab_alloc(h, 4 * G);
  uint32_t* d_cn = (uint32_t*)slab_alloc(h, 4 * (size_t)n_const + 4);
  uint32_t* d_cv = (uint32_t*)slab_alloc(h, 4 * (size_t)n_const + 4);
  uint32_t* d_on = (uint32_t*)slab_alloc(h, 4 * (size_t)n_out + 4);
  if (!ok || !d_gates || !d_lo || !d_off || !node_const || !out_mark || !node_val || !d_cmask || !d_live || !d_dead || !d_cval || !d_cn || !d_cv || !d_on)
    return fail(h, C2A_ERR_NO_MEMORY, "scratch slab exhausted");
  cudaStream_t s = h->stream;
  if (G) cudaMemcpyAsync(d_gates, gates, 16 * G, cudaMemcpyHostToDevice, s);
  if (n_const) { cudaMemcpyAsync(d_cn, const_nodes, 4 * (size_t)n_const, cudaMemcpyHostToDevice, s); cudaMemcpyAsync(d_cv, const_values, 4 * (size_t)n_const, cudaMemcpyHostToDevice, s); }
  if (n_out) cudaMemcpyAsync(d_on, output_nodes, 4 * (size_t)n_out, cudaMemcpyHostToDevice, s);
  cudaMemsetAsync(d_off, 0, 4 * (G + 2), s);
  uint32_t nl = 0;
  st = kahn_core(h, d_gates, (uint32_t)G, node_bound, b, d_lo, d_off, (uint32_t)std::min<uint64_t>(G, 0xFFFFFFFEull), &nl, err_index);
  if (st != C2A_OK) { phases_collect(h); return st; }
  std::vector<uint32_t> off((size_t)nl + 1, 0);
  cudaMemcpyAsync(off.data(), d_off, 4 * ((size_t)nl + 1), cudaMemcpyDeviceToHost, s);
  cudaMemsetAsync(node_const, 0, node_bound, s);
  cudaMemsetAsync(out_mark, 0, node_bound, s);
  cudaMemsetAsync(node_val, 0, 4 * (size_t)node_bound, s);
  if (!cuda_ok(h, cudaStreamSynchronize(s), "sweep setup")) return C2A_ERR_CUDA;
  off[nl] = (uint32_t)G;
  if (n_const) LAUNCH(h, k_set_node_consts, grid_for(h, (const void*)k_set_node_consts, kBlock, n_const), kBlock, d_cn, d_cv, n_const, node_bound, node_const, node_val);
  if (n_out) LAUNCH(h, k_mark_nodes, grid_for(h, (const void*)k_mark_nodes, kBlock, n_out), kBlock, d_on, n_out, node_bound, out_mark);
  phase_begin(h, "k_fold_level");
  for (uint32_t l = 0; l < nl; ++l)
    if (off[l + 1] > off[l]) LAUNCH(h, k_fold_level, grid_for(h, (const void*)k_fold_level, kBlock, off[l + 1] - off[l]), kBlock, d_gates, d_lo, off[l], off[l + 1], b.prod1, node_const, node_val, d_cmask, d_cval);
  phase_end(h);
  phase_begin(h, "k_live_level");
  for (uint32_t l = nl; l-- > 0;)
    if (off[l + 1] > off[l]) LAUNCH(h, k_live_level, grid_for(h, (const void*)k_live_level, kBlock, off[l + 1] - off[l]), kBlock, d_gates, d_lo, off[l], off[l + 1], b.prod1, b.row_off, b.col, out_mark, d_live);
  if (G) LAUNCH(h, k_invert_u8, grid_for(h, (const void*)k_invert_u8, kBlock, G), kBlock, d_live, d_dead, (uint32_t)G);
  phase_end(h);
  if (G) {
    if (const_mask) cudaMemcpyAsync(const_mask, d_cmask, G, cudaMemcpyDeviceToHost, s);
    if (const_value) cudaMemcpyAsync(const_value, d_cval, 4 * G, cudaMemcpyDeviceToHost, s);
    if (dead_mask) cudaMemcpyAsync(dead_mask, d_dead, G, cudaMemcpyDeviceToHost, s);
  }
  if (!cuda_ok(h, cudaStreamSynchronize(s), "sweep D2H")) return C2A_ERR_CUDA;
  if (!cuda_ok(h, cudaGetLastError(), "sweep kernels")) return C2A_ERR_CUDA;
  phases_collect(h);
  return C2A_OK;
}

}  // extern "C"
