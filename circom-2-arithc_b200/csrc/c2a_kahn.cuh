// c2a_kahn.cuh — K3/K4: consumer CSR + Kahn levels, and the opt-in layer-wise sweeps (K8 constant-fold mask, K9 dead-gate
// mask).  Included by c2a_device.cu (single translation unit).
//
// The dependency relation is the reference's (src/compiler.rs:401-421): gate g depends on the LAST producer
// of its lh and rh nodes.  Kahn yields a valid topological order and the levels (longest-path depth), NOT the reference's
// DFS post-order (that is K5, c2a_device.cu); it is used for cycle screening, the sweeps and the evaluator.
//
//   K3  (row lengths counted by k_deps_t<true>) / k_scan_u32 / k_kahn_fill   consumer CSR  row_off[G+1], col[<=2G].  A col entry is 16 bytes:
//       {consumer | two-producer flag << 31, the consumer's own row begin, its row length, 0} - the walker that releases the
//       consumer already holds its row, so a hop costs one dependent load (plus one atomic for a two-producer consumer).
//       A gate that reads the same producer on both operands is listed once and counts as a one-producer consumer.
//   K4  asynchronous Kahn, no level barrier (round 1 measured the level-synchronous cooperative kernel at 8.5 us per level -
//       one grid barrier + a six-deep dependent load/atomic chain - which is 4.7 ms on the 547-level, 10 M-gate MiMC stream):
//         k_kahn_walk     every gate without dependencies is a root; the thread that releases a consumer keeps walking it
//                         (chain following: ~2-3 dependent L2 round trips per hop, no synchronisation), further released
//                         consumers go to a small per-thread stack, its overflow to a global queue (next launch).
//                         A one-slot consumer is released by its only producer with level L+1, no atomic at all;
//                         a two-slot consumer by ONE atomicMax(lv[c], L+1): the first arrival reads 0, the second reads the
//                         other level - in-degree decrement and level maximum in a single word.
//         k_kahn_long     rows of >= kLongRow consumers: one CTA per row, the row staged through shared memory by TMA bulk
//                         copies (cp.async.bulk + mbarrier), released consumers appended with one warp-aggregated atomicAdd
//         k_level_pass<hist> / scan / k_level_pass<scatter>   counting sort by level -> level-major order + level_off[]
#pragma once

namespace c2a {

constexpr int kKahnBlock = 512;
constexpr uint32_t kLongRow = 512;        // consumers; rows at least this long are walked by a whole CTA
constexpr int kWalkStack = 16;            // released-but-not-yet-walked consumers a thread keeps for itself
constexpr uint32_t kTwoSlots = 0x80000000u;
enum { KC_QN0 = 0, KC_QN1 = 1, KC_LONGN0 = 2, KC_LONGN1 = 3, KC_DONE = 4, KC_MAXLV = 5, KC_ERRMIN = 6, KC_COUNT = 16 };

__global__ void __launch_bounds__(kBlock) k_kahn_fill(const uint2* __restrict__ dep, uint32_t G, const uint32_t* __restrict__ row_off,
                                                      uint32_t* __restrict__ cursor, uint4* __restrict__ col) {
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < G; g += gridDim.x * kBlock) {
    uint2 d = dep[g];
    if (d.y == d.x) d.y = kNone;
    if (d.x == kNone && d.y == kNone) continue;
    const uint32_t beg = row_off[g], len = row_off[g + 1] - beg;
    const uint4 e = make_uint4(g | ((d.x != kNone && d.y != kNone) ? kTwoSlots : 0u), beg, len, 0u);
    if (d.x != kNone) col[row_off[d.x] + atomicAdd(cursor + d.x, 1u)] = e;
    if (d.y != kNone) col[row_off[d.y] + atomicAdd(cursor + d.y, 1u)] = e;
  }
}

// warp-aggregated append of `item` for lanes with ready != 0 (all 32 lanes must call)
template <typename T>
__device__ __forceinline__ void warp_append(bool ready, const T& item, T* __restrict__ queue, uint32_t* __restrict__ tail) {
  uint32_t m = __ballot_sync(0xFFFFFFFFu, ready);
  if (!m) return;
  int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(tail, (uint32_t)__popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
  if (ready) queue[base + __popc(m & ((1u << lane) - 1))] = item;
}

// one arrival at consumer entry e from a producer of level L: returns true when this arrival releases it (*lc = its level)
__device__ __forceinline__ bool kahn_arrive(uint32_t e, uint32_t L, uint32_t* __restrict__ lv, uint32_t* lc) {
  if (!(e & kTwoSlots)) { *lc = L + 1; return true; }
  uint32_t old = atomicMax(lv + (e & ~kTwoSlots), L + 1);  // lv = 0: nobody arrived yet (levels stored here are >= 1)
  *lc = max(old, L + 1);
  return old != 0;
}

// Walk from gate g (level L, consumer row [beg, beg+len)): record the level, release consumers, continue with one of them.
// queue entries are {gate, level, row begin, row length}.
__device__ __forceinline__ void kahn_walk(uint32_t g, uint32_t L, uint32_t beg, uint32_t len, const uint4* __restrict__ col,
                                          uint32_t* __restrict__ lv, uint32_t* __restrict__ level_of, uint4* __restrict__ q_out,
                                          uint32_t* __restrict__ q_out_n, uint4* __restrict__ long_out, uint32_t* __restrict__ long_out_n,
                                          uint32_t& done, uint32_t& maxl) {
  uint4 stack[kWalkStack];
  int sp = 0;
  while (true) {
    level_of[g] = L;
    ++done;
    maxl = max(maxl, L);
    uint4 nxt = make_uint4(kNone, 0, 0, 0);
    if (len >= kLongRow) {
      long_out[atomicAdd(long_out_n, 1u)] = make_uint4(g, L, beg, len);  // a whole CTA takes this row (k_kahn_long)
    } else {
      for (uint32_t j = 0; j < len; ++j) {
        uint4 e = __ldg(col + beg + j);
        uint32_t lc;
        if (!kahn_arrive(e.x, L, lv, &lc)) continue;
        uint4 item = make_uint4(e.x & ~kTwoSlots, lc, e.y, e.z);
        if (nxt.x == kNone) nxt = item;
        else if (sp < kWalkStack) stack[sp++] = item;
        else q_out[atomicAdd(q_out_n, 1u)] = item;
      }
    }
    if (nxt.x == kNone) {
      if (sp == 0) break;
      nxt = stack[--sp];
    }
    g = nxt.x; L = nxt.y; beg = nxt.z; len = nxt.w;
  }
}

__device__ __forceinline__ void kahn_commit(uint32_t done, uint32_t maxl, uint32_t* ctrl) {
  done = warp_sum(done);
  maxl = warp_max(maxl);
  if ((threadIdx.x & 31) == 0 && done) { atomicAdd(ctrl + KC_DONE, done); atomicMax(ctrl + KC_MAXLV, maxl); }
}

// roots: gates without dependencies (level 0)
__global__ void __launch_bounds__(kBlock) k_kahn_walk_roots(const uint2* __restrict__ dep, uint32_t G, const uint32_t* __restrict__ row_off,
                                                            const uint4* __restrict__ col, uint32_t* __restrict__ lv, uint32_t* __restrict__ level_of,
                                                            uint4* __restrict__ q_out, uint4* __restrict__ long_out, uint32_t* ctrl) {
  uint32_t done = 0, maxl = 0;
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < G; g += gridDim.x * kBlock) {
    uint2 d = dep[g];
    if (d.x == kNone && d.y == kNone) {
      uint32_t beg = row_off[g];
      kahn_walk(g, 0, beg, row_off[g + 1] - beg, col, lv, level_of, q_out, ctrl + KC_QN0, long_out, ctrl + KC_LONGN0, done, maxl);
    }
  }
  kahn_commit(done, maxl, ctrl);
}

// overflow queue of the previous launch: released consumers nobody has walked yet
__global__ void __launch_bounds__(kBlock) k_kahn_walk_queue(const uint4* __restrict__ q_in, const uint32_t* __restrict__ q_in_n,
                                                            const uint4* __restrict__ col, uint32_t* __restrict__ lv, uint32_t* __restrict__ level_of,
                                                            uint4* __restrict__ q_out, uint32_t* __restrict__ q_out_n, uint4* __restrict__ long_out,
                                                            uint32_t* __restrict__ long_out_n, uint32_t* ctrl) {
  uint32_t done = 0, maxl = 0;
  const uint32_t n = *q_in_n;
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    uint4 it = q_in[i];
    kahn_walk(it.x, it.y, it.z, it.w, col, lv, level_of, q_out, q_out_n, long_out, long_out_n, done, maxl);
  }
  kahn_commit(done, maxl, ctrl);
}

// TMA 1-D bulk copy global -> shared, completion on an mbarrier (cp.async.bulk; SASS: UBLKCP)
__device__ __forceinline__ void bulk_load_row(uint32_t* smem_dst, const uint32_t* gsrc, uint32_t bytes, unsigned long long* mbar, uint32_t phase) {
  uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
  }
  // all threads wait for the phase to complete
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
    if (!done && ++spins > (1u << 24)) __trap();  // a lost completion surfaces as a CUDA error, not a hang
  }
}

// long rows: one CTA per row; the gate itself was already recorded by the walker that found it.  The row (16-byte entries, so
// every chunk is bulk-copy aligned) is staged through shared memory; released consumers are queued with their own row
// (they are walked by the next k_kahn_walk_queue launch).
constexpr uint32_t kBulkEntries = 1024;  // 16 KB per bulk copy
__global__ void __launch_bounds__(kKahnBlock) k_kahn_long(const uint4* __restrict__ long_in, const uint32_t* __restrict__ long_in_n,
                                                          const uint4* __restrict__ col, uint32_t* __restrict__ lv, uint4* __restrict__ q_out,
                                                          uint32_t* __restrict__ q_out_n) {
  __shared__ __align__(16) uint4 s_stage[kBulkEntries];
  __shared__ __align__(8) unsigned long long s_mbar;
  uint32_t mbar_phase = 0;
  if (threadIdx.x == 0) {
    uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t n = *long_in_n;
  for (uint32_t r = blockIdx.x; r < n; r += gridDim.x) {
    const uint4 it = long_in[r];  // {gate, level, row begin, row length}
    const uint32_t L = it.y;
    for (uint32_t pos = 0; pos < it.w; pos += kBulkEntries) {
      const uint32_t cnt = min(it.w - pos, kBulkEntries);
      bulk_load_row(reinterpret_cast<uint32_t*>(s_stage), reinterpret_cast<const uint32_t*>(col + it.z + pos), cnt * 16, &s_mbar, mbar_phase);
      mbar_phase ^= 1;
      for (uint32_t j = threadIdx.x; j < ((cnt + 31) / 32) * 32; j += blockDim.x) {
        bool v = j < cnt;
        uint4 e = v ? s_stage[j] : make_uint4(0, 0, 0, 0);
        uint32_t lc = 0;
        bool ready = v && kahn_arrive(e.x, L, lv, &lc);
        warp_append(ready, make_uint4(e.x & ~kTwoSlots, lc, e.y, e.z), q_out, q_out_n);
      }
      __syncthreads();  // everyone is done with the staging buffer
    }
  }
}

// Counting sort by level.  A 4096-gate tile of the (chain-major) gate vector spans few distinct levels, so both passes count in
// shared memory and touch the global per-level counters once per (tile, level) instead of once per gate; tiles spanning more
// than kLvTile levels fall back to per-gate global atomics.  kScatter = false: histogram; true: reserve + place.
constexpr int kLvItems = 16, kLvTile = kBlock * kLvItems;
template <bool kScatter>
__global__ void __launch_bounds__(kBlock) k_level_pass(const uint32_t* __restrict__ level_of, uint32_t G, uint32_t* __restrict__ counter,
                                                       const uint32_t* __restrict__ level_off, uint32_t* __restrict__ level_order) {
  __shared__ uint32_t s_bin[kLvTile];
  __shared__ uint32_t s_min, s_max;
  const uint32_t base = blockIdx.x * kLvTile;
  uint32_t lv[kLvItems], rk[kLvItems];
  uint32_t mn = 0xFFFFFFFFu, mx = 0;
  if (threadIdx.x == 0) { s_min = 0xFFFFFFFFu; s_max = 0; }
#pragma unroll
  for (int i = 0; i < kLvItems; ++i) {
    uint32_t g = base + i * kBlock + threadIdx.x;
    lv[i] = g < G ? level_of[g] : 0xFFFFFFFFu;
    if (g < G) { mn = min(mn, lv[i]); mx = max(mx, lv[i]); }
  }
  __syncthreads();
#pragma unroll
  for (int o = 16; o; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o)); mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o)); }
  if ((threadIdx.x & 31) == 0 && mn != 0xFFFFFFFFu) { atomicMin(&s_min, mn); atomicMax(&s_max, mx); }
  __syncthreads();
  const uint32_t lo = s_min, range = s_max - lo + 1;
  if (range <= (uint32_t)kLvTile) {
    for (uint32_t b = threadIdx.x; b < range; b += kBlock) s_bin[b] = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kLvItems; ++i)
      if (lv[i] != 0xFFFFFFFFu) rk[i] = atomicAdd(&s_bin[lv[i] - lo], 1u);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < range; b += kBlock) {
      uint32_t c = s_bin[b];
      if (c) {
        uint32_t at = atomicAdd(counter + lo + b, c);
        if (kScatter) s_bin[b] = level_off[lo + b] + at;  // where this tile's members of the level go
      }
    }
    if (kScatter) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kLvItems; ++i)
        if (lv[i] != 0xFFFFFFFFu) level_order[s_bin[lv[i] - lo] + rk[i]] = base + i * kBlock + threadIdx.x;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kLvItems; ++i)
      if (lv[i] != 0xFFFFFFFFu) {
        uint32_t at = atomicAdd(counter + lv[i], 1u);
        if (kScatter) level_order[level_off[lv[i]] + at] = base + i * kBlock + threadIdx.x;
      }
  }
}

// after the walk ran dry: gates that were never released sit on or behind a cycle
__global__ void __launch_bounds__(kBlock) k_kahn_leftover(const uint32_t* __restrict__ level_of, uint32_t G, uint32_t* ctrl) {
  uint32_t m = kNone;
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < G; g += gridDim.x * kBlock)
    if (level_of[g] == kNone) { m = g; break; }
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

// node_const[node]: bit0 = value known; node_val[node] = value (k_sweep_levels<true> below).
__global__ void __launch_bounds__(kBlock) k_set_node_consts(const uint32_t* __restrict__ nodes, const uint32_t* __restrict__ vals, uint32_t n,
                                                            uint32_t node_bound, uint8_t* __restrict__ node_const, uint32_t* __restrict__ node_val) {
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock)
    if (nodes[i] < node_bound) { node_const[nodes[i]] = 1; node_val[nodes[i]] = vals[i]; }
}

__global__ void __launch_bounds__(kBlock) k_mark_nodes(const uint32_t* __restrict__ nodes, uint32_t n, uint32_t node_bound, uint8_t* __restrict__ mark) {
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock)
    if (nodes[i] < node_bound) mark[nodes[i]] = 1;
}

// (reverse sweep: a gate is live when it produces an output node or feeds a live gate)
// Both sweeps as ONE cooperative launch each: the levels are walked on the device with a grid barrier in between (a launch per
// level - 547 on the MiMC circuit, thousands on a bit-sliced SHA-256 - costs ~6 us of launch latency per level; the barrier ~2 us).
// Values written for level l are read by other CTAs in level l +- 1: those reads go through L2 (ld.cg).
template <bool kFold>
__global__ void __launch_bounds__(kBlock) k_sweep_levels(const uint4* __restrict__ gates, const uint32_t* __restrict__ level_order,
                                                         const uint32_t* __restrict__ level_off, uint32_t nl, uint32_t G, const uint32_t* __restrict__ prod1,
                                                         uint8_t* node_const, uint32_t* node_val, uint8_t* __restrict__ const_mask, uint32_t* __restrict__ const_value,
                                                         const uint32_t* __restrict__ row_off, const uint4* __restrict__ col, const uint8_t* __restrict__ out_mark,
                                                         uint8_t* live, unsigned int* bar) {
  unsigned int epoch = 0;
  const uint32_t tid = blockIdx.x * kBlock + threadIdx.x;
  auto range = [&](uint32_t step, uint32_t& lo, uint32_t& hi) {
    const uint32_t l = kFold ? step : nl - 1 - step;
    lo = __ldg(level_off + l);
    hi = l + 1 == nl ? G : __ldg(level_off + l + 1);
  };
  // a level is a chain of dependent loads (order -> gate -> operand values): the first two do not depend on the previous level and
  // are fetched BEFORE the barrier that ends it
  uint32_t lo, hi, g_pre = 0;
  uint4 gt_pre = make_uint4(0, 0, 0, 0);
  if (nl) { range(0, lo, hi); if (lo + tid < hi) { g_pre = __ldg(level_order + lo + tid); gt_pre = __ldg(gates + g_pre); } }
  for (uint32_t step = 0; step < nl; ++step) {
    for (uint32_t i = lo + tid; i < hi; i += gridDim.x * kBlock) {
      const bool first = i == lo + tid;
      const uint32_t g = first ? g_pre : __ldg(level_order + i);
      const uint4 gt = first ? gt_pre : __ldg(gates + g);
      if (kFold) {
        uint32_t r = 0;
        const bool c = __ldcg(node_const + gt.y) && __ldcg(node_const + gt.z) && exec_op_u32(gt.x, __ldcg(node_val + gt.y), __ldcg(node_val + gt.z), &r);
        const_mask[g] = c;
        const_value[g] = c ? r : 0u;
        if (__ldg(prod1 + gt.w) == g + 1) {  // this gate is the node's producer (last writer wins, compiler.rs:403-406)
          node_const[gt.w] = c;
          node_val[gt.w] = r;
        }
      } else {
        bool lv = __ldg(out_mark + gt.w) && __ldg(prod1 + gt.w) == g + 1;
        const uint32_t e = __ldg(row_off + g + 1);
        for (uint32_t j = __ldg(row_off + g); !lv && j < e; ++j) lv = __ldcg(live + (__ldg(col + j).x & ~kTwoSlots)) != 0;
        live[g] = lv;
      }
    }
    if (step + 1 < nl) {
      range(step + 1, lo, hi);
      if (lo + tid < hi) { g_pre = __ldg(level_order + lo + tid); gt_pre = __ldg(gates + g_pre); }
      sort_grid_bar(bar, epoch);
    }
  }
}

__global__ void __launch_bounds__(kBlock) k_invert_u8(const uint8_t* __restrict__ a, uint8_t* __restrict__ b, uint32_t n) {
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) b[i] = a[i] ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------------
struct KahnBuffers {
  uint32_t* prod1;
  uint2* dep;
  uint32_t* row_off;   // G+1
  uint32_t* cursor;    // G: CSR fill cursors, later the per-level scatter cursors
  uint4* col;          // 2G entries of 16 bytes
  uint32_t* lv;        // G: arrival word of two-slot consumers
  uint32_t* level_of;  // G
  uint4* q[2];         // overflow queues {gate, level, row begin, row length}
  uint4* longq[2];     // long-row queues, same record
  unsigned long long* tile_state;
  uint32_t* scalars;
  uint32_t* ctrl;
};

static inline size_t kahn_long_cap(uint64_t G) { return (size_t)(G / (kLongRow / 2)) + 64; }  // <= 2G consumer entries in total

static size_t kahn_scratch_bytes(uint64_t G, uint32_t node_bound) {
  return align256(4 * (size_t)node_bound) + align256(8 * G) + align256(4 * (G + 1)) + 3 * align256(4 * G) + align256(32 * G + 16) + 2 * align256(16 * G) +
         2 * align256(16 * kahn_long_cap(G)) + align256(8 * (size_t)(scan_tiles(G + 1, kScanItems) + 1)) + align256(4 * S_COUNT) + align256(4 * KC_COUNT);
}

static bool kahn_carve(c2a_handle* h, uint64_t G, uint32_t node_bound, KahnBuffers* b) {
  b->prod1 = (uint32_t*)slab_alloc(h, 4 * (size_t)node_bound);
  b->dep = (uint2*)slab_alloc(h, 8 * G);
  b->row_off = (uint32_t*)slab_alloc(h, 4 * (G + 1));
  b->cursor = (uint32_t*)slab_alloc(h, 4 * G);
  b->col = (uint4*)slab_alloc(h, 32 * G + 16);
  b->lv = (uint32_t*)slab_alloc(h, 4 * G);
  b->level_of = (uint32_t*)slab_alloc(h, 4 * G);
  for (int i = 0; i < 2; ++i) b->q[i] = (uint4*)slab_alloc(h, 16 * G);
  for (int i = 0; i < 2; ++i) b->longq[i] = (uint4*)slab_alloc(h, 16 * kahn_long_cap(G));
  b->tile_state = (unsigned long long*)slab_alloc(h, 8 * (size_t)(scan_tiles(G + 1, kScanItems) + 1));
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
  hp[S_COUNT + KC_ERRMIN] = kNone;
  cudaMemcpyAsync(b.scalars, hp, 4 * S_COUNT, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(b.ctrl, hp + S_COUNT, 4 * KC_COUNT, cudaMemcpyHostToDevice, st);
  phase_begin(h, "init");
  cudaMemsetAsync(b.prod1, 0, 4 * (size_t)node_bound, st);
  cudaMemsetAsync(b.row_off, 0, 4 * ((size_t)G + 1), st);
  cudaMemsetAsync(b.cursor, 0, 4 * (size_t)G, st);
  cudaMemsetAsync(b.lv, 0, 4 * (size_t)G, st);
  cudaMemsetAsync(b.level_of, 0xFF, 4 * (size_t)G, st);
  phase_end(h);
  uint32_t* hc = hp + S_COUNT;  // host copy of ctrl
  if (G) {
    phase_begin(h, "k_producer");
    LAUNCH(h, k_producer, grid_for(h, (const void*)k_producer, kBlock, G), kBlock, d_gates, G, node_bound, b.prod1, b.scalars);
    phase_end(h);
    phase_begin(h, "k_deps");
    // K2 + the row lengths of the consumer CSR in one pass (k_kahn_count fused into the dependency kernel)
    LAUNCH(h, k_deps_t<true>, grid_for(h, (const void*)k_deps_t<true>, kBlock, G), kBlock, d_gates, G, node_bound, b.prod1, b.dep, (uint32_t*)nullptr, b.row_off, b.scalars);
    phase_end(h);
    uint32_t tiles = scan_tiles(G, kScanItems);
    cudaMemsetAsync(b.tile_state, 0, 8 * (size_t)tiles, st);
    phase_begin(h, "k_scan_u32");
    LAUNCH(h, k_scan_u32_t<false>, tiles, kBlock, b.row_off, b.row_off, G, b.tile_state, (uint32_t*)nullptr, (const uint32_t*)nullptr, 0);
    phase_end(h);
    phase_begin(h, "k_kahn_fill");
    LAUNCH(h, k_kahn_fill, grid_for(h, (const void*)k_kahn_fill, kBlock, G), kBlock, b.dep, G, b.row_off, b.cursor, b.col);
    phase_end(h);
    // ---- K4: walk from the roots; overflow / long rows are handed to follow-up launches until nothing is queued
    phase_begin(h, "k_kahn_walk");
    LAUNCH(h, k_kahn_walk_roots, grid_for(h, (const void*)k_kahn_walk_roots, kBlock, G), kBlock, b.dep, G, b.row_off, b.col, b.lv, b.level_of, b.q[0], b.longq[0], b.ctrl);
    phase_end(h);
    int cur = 0;
    for (int round = 0;; ++round) {
      cudaMemcpyAsync(hc, b.ctrl, 4 * KC_COUNT, cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(hp, b.scalars, 4 * S_COUNT, cudaMemcpyDeviceToHost, st);
      if (!cuda_ok(h, cudaStreamSynchronize(st), "kahn sync")) return C2A_ERR_CUDA;
      if (!cuda_ok(h, cudaGetLastError(), "kahn kernels")) return C2A_ERR_CUDA;
      if (hp[S_FLAGS] & F_BAD) return fail(h, C2A_ERR_INVALID_ARGUMENT, "a gate references a node id >= node_bound (%u)", node_bound);
      const uint32_t nq = hc[KC_QN0 + cur], nlong = hc[KC_LONGN0 + cur];
      if (!nq && !nlong) break;
      if (round > (int)G + 8) return fail(h, C2A_ERR_CUDA, "Kahn walk did not converge");
      const int nxt = cur ^ 1;
      cudaMemsetAsync(b.ctrl + KC_QN0 + nxt, 0, 4, st);
      cudaMemsetAsync(b.ctrl + KC_LONGN0 + nxt, 0, 4, st);
      if (nlong) {
        phase_begin(h, "k_kahn_long");
        LAUNCH(h, k_kahn_long, std::min<uint32_t>(nlong, (uint32_t)h->num_sms * 2), kKahnBlock, b.longq[cur], b.ctrl + KC_LONGN0 + cur, b.col, b.lv, b.q[nxt],
               b.ctrl + KC_QN0 + nxt);
        phase_end(h);
      }
      if (nq) {
        phase_begin(h, "k_kahn_walk");
        LAUNCH(h, k_kahn_walk_queue, grid_for(h, (const void*)k_kahn_walk_queue, kBlock, nq), kBlock, b.q[cur], b.ctrl + KC_QN0 + cur, b.col, b.lv, b.level_of,
               b.q[nxt], b.ctrl + KC_QN0 + nxt, b.longq[nxt], b.ctrl + KC_LONGN0 + nxt, b.ctrl);
        phase_end(h);
      }
      cur = nxt;
    }
    if (hc[KC_DONE] != G) {  // gates that were never released sit on or behind a cycle
      LAUNCH(h, k_kahn_leftover, grid_for(h, (const void*)k_kahn_leftover, kBlock, G), kBlock, b.level_of, G, b.ctrl);
      cudaMemcpyAsync(hc, b.ctrl, 4 * KC_COUNT, cudaMemcpyDeviceToHost, st);
      if (!cuda_ok(h, cudaStreamSynchronize(st), "kahn leftover")) return C2A_ERR_CUDA;
      if (err_index) *err_index = hc[KC_ERRMIN];
      return fail(h, C2A_ERR_CYCLIC_DEPENDENCY, "%u of %u gates are on or behind a dependency cycle (smallest index %u)", G - hc[KC_DONE], G, hc[KC_ERRMIN]);
    }
  } else {
    if (!cuda_ok(h, cudaStreamSynchronize(st), "kahn sync")) return C2A_ERR_CUDA;
  }
  const uint32_t levels = G ? hc[KC_MAXLV] + 1 : 0;
  if (levels > level_cap) return fail(h, C2A_ERR_INVALID_ARGUMENT, "more levels than level_cap (%u)", level_cap);
  // ---- counting sort by level -> level-major order; level_off = exclusive scan of the level sizes (level_off[levels] = G)
  cudaMemsetAsync(d_level_off, 0, 4 * ((size_t)levels + 1), st);
  if (G) {
    phase_begin(h, "k_level_sort");
    cudaMemsetAsync(b.cursor, 0, 4 * (size_t)levels, st);
    const uint32_t lvtiles = (G + kLvTile - 1) / kLvTile;
    LAUNCH(h, k_level_pass<false>, lvtiles, kBlock, b.level_of, G, d_level_off, (const uint32_t*)nullptr, (uint32_t*)nullptr);
    uint32_t ltiles = scan_tiles(levels, kScanItems);
    cudaMemsetAsync(b.tile_state, 0, 8 * (size_t)ltiles, st);
    LAUNCH(h, k_scan_u32_t<false>, ltiles, kBlock, d_level_off, d_level_off, levels, b.tile_state, (uint32_t*)nullptr, (const uint32_t*)nullptr, 0);
    LAUNCH(h, k_level_pass<true>, lvtiles, kBlock, b.level_of, G, b.cursor, d_level_off, d_level_order);
    phase_end(h);
  }
  if (!cuda_ok(h, cudaStreamSynchronize(st), "kahn level sort")) return C2A_ERR_CUDA;
  if (!cuda_ok(h, cudaGetLastError(), "kahn level sort kernels")) return C2A_ERR_CUDA;
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
  // one cooperative launch per sweep (the barrier counters live in the Kahn control words, unused by now)
  cudaMemsetAsync(b.ctrl, 0, 4 * KC_COUNT, s);
  auto sweep = [&](bool fold) {
    const uint4* a_gates = d_gates;
    const uint32_t *a_lo = d_lo, *a_off = d_off, *a_prod = b.prod1, *a_row = b.row_off;
    uint32_t a_nl = nl, a_G = (uint32_t)G;
    uint8_t *a_nc = node_const, *a_cm = d_cmask, *a_live = d_live;
    const uint8_t* a_om = out_mark;
    uint32_t *a_nv = node_val, *a_cv = d_cval;
    const uint4* a_col = b.col;
    unsigned int* a_bar = reinterpret_cast<unsigned int*>(b.ctrl) + (fold ? 0 : 1);
    void* args[] = {&a_gates, &a_lo, &a_off, &a_nl, &a_G, &a_prod, &a_nc, &a_nv, &a_cm, &a_cv, &a_row, &a_col, &a_om, &a_live, &a_bar};
    const void* fn = fold ? (const void*)k_sweep_levels<true> : (const void*)k_sweep_levels<false>;
    cudaLaunchCooperativeKernel(fn, dim3(h->num_sms), dim3(kBlock), args, 0, s);
    h->launches++;
  };
  phase_begin(h, "k_fold_level");
  if (nl) sweep(true);
  phase_end(h);
  phase_begin(h, "k_live_level");
  if (nl) sweep(false);
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
