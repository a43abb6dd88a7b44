// c2a_device.cu — sm_100a kernels + device half of the C ABI (include/c2a.h).
//
// Replaces, bit-for-bit, the back end of the reference's Compiler::build_circuit
// (src/compiler.rs:388-464) and its DFS topological sort (src/topological_sort.rs:3-50).
//
// Pipeline (all integer/index work, HBM/L2-bound, no tensor cores):
//   K1 producer_map   prod1[out[g]] = max(g+1)                       compiler.rs:401-406 (last insert wins)
//   K2 deps           dep[g] = {prod1[lh]-1, prod1[rh]-1}, out-of-order edge detection
//                     (no dep >= g  <=>  the DFS post-order is 0..G-1)  compiler.rs:408-421
//   K5a relax         r[v] = min(v, min over consumers r[u]) = the DFS root that first reaches v
//   K5b sizes/scan    block offsets per root (roots ascending)          topological_sort.rs:11-13
//   K5c trees         per-root DFS post-order restricted to its block, lh before rh, cycle detection
//                                                                       topological_sort.rs:23-50
//   K6 wire_first/wire_scan   first-seen wire numbering                 compiler.rs:427-443
//   K7 gather         new_gates[k] = {op, wire[lh], wire[rh], wire[out]}  compiler.rs:452-464
//   K4 kahn           level-synchronous frontier (levels for the sweeps; c2a_kahn.cu)
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only; ranges cost a pointer test unless a tool (nsys / ncu --nvtx) is attached

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "c2a_internal.h"

namespace c2a {

// ---------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  // 128-bit load of a gate record; no L1 allocation (each pass touches a record once), L2 keeps it
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
// look-back tile states: relaxed, GPU scope (ld/st.volatile compile to .STRONG.SYS accesses)
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// warp reductions
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t warp_max(uint32_t v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t warp_or(uint32_t v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v |= __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

constexpr int kBlock = 256;
constexpr uint32_t kNone = C2A_NONE;
constexpr uint32_t kBigBlock = 1024;  // DFS blocks at least this large go through the pointer-jumping path when they are trees
static void emit_drop_host(c2a_handle* h);  // c2a_emit.cuh
// wire[] encoding while numbering is in flight (compiler.rs:388-449):
//   < kOutPending        : final wire id
//   == kOutPending       : output node, numbered after all intermediates (:431-434, :446-449)
//   kFirstTag | p        : not numbered yet; p = 3*pos+slot of the earliest appearance seen so far
//   kNone                : never appears
constexpr uint32_t kOutPending = 0x7FFFFFFFu;
constexpr uint32_t kFirstTag = 0x80000000u;

// scalars[] slots on the device
// S_QN0 .. S_QN0+kSpecRounds: queue lengths of the speculative relax rounds (round j reads S_QN0+j, appends to S_QN0+j+1);
// S_QN_LAST != 0 after them means "r[] has not converged yet": every later kernel of the call parks itself and the host
// drains the queue with synchronised rounds (S_QA / S_QB) before re-issuing the tail of the pipeline.
// S_QN0 .. S_QN0+3: rotating queue lengths of the relaxation rounds (k_relax_loop); S_BAR*: grid-barrier counters of the two
// cooperative kernels of the sort; S_RELAX_ROUNDS / S_ROBUST: diagnostics (rounds run, pointer-jumping stage taken);
// S_CHG0..+2: "something changed" words of the jumping rounds (round i sets word i % 3 and clears word (i + 1) % 3, which was last
// read two barriers ago); S_BIGN: nodes in big tree-shaped DFS blocks (k_tree_blocks)
enum { S_FLAGS = 0, S_HEAVYN = 1, S_NMID = 2, S_SEEDN = 3, S_ERR_LO = 6, S_ERR_HI = 7, S_QN0 = 8,
       S_BAR1 = 16, S_BAR2 = 17, S_RELAX_ROUNDS = 18, S_ROBUST = 19, S_CHG0 = 20 /* 3 rotating words */, S_BIGN = 23, S_BIGBLK = 24, S_CAP0 = 25 /* 3 rotating: walks cut at kHopCap per round */, S_COUNT = 32 };
enum { F_OOO = 1, F_BAD = 2, F_SELF = 4, F_BAD_IO = 8 };

// Device-side control flow: the host enqueues the whole pipeline without looking at intermediate results; kernels decide
// from the scalars whether they have work.  (One status read at the very end instead of a round trip per decision.)
__device__ __forceinline__ bool sort_wanted(const uint32_t* __restrict__ sc) {  // the DFS order differs from 0..n-1
  uint32_t f = sc[S_FLAGS];
  return !(f & F_BAD) && (f & (F_OOO | F_SELF));
}
__device__ __forceinline__ bool sort_parked(const uint32_t*) { return false; }  // (the relaxation always converges on the device: k_relax_loop)
__device__ __forceinline__ bool tail_parked(const uint32_t* __restrict__ sc) {  // kernels after the sort: bad input, cycle found
  return (sc[S_FLAGS] & F_BAD) || sc[S_ERR_HI] != 0xFFFFFFFFu;
}

// ---------------------------------------------------------------------------------------------------
// K1: producer map.  prod1[node] = 1 + (largest gate index whose out is node); 0 = no producer.
// compiler.rs:403-406: HashMap::insert in ascending gate order => the LAST gate wins.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_producer(const uint4* __restrict__ gates, uint32_t G, uint32_t node_bound,
                                                     uint32_t* __restrict__ prod1, uint32_t* __restrict__ scalars) {
  bool bad = false;
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < G; g += gridDim.x * kBlock) {
    uint4 gt = ldg_stream(gates + g);
    if (gt.y >= node_bound || gt.z >= node_bound || gt.w >= node_bound) bad = true;
    else atomicMax(prod1 + gt.w, g + 1);  // RED.MAX, no return value
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(scalars + S_FLAGS, (uint32_t)F_BAD);
}

// ---------------------------------------------------------------------------------------------------
// K2: dependencies.  dep[g] = (producer of lh, producer of rh), kNone where the node has no producer
// (compiler.rs:412-418).  flags |= F_OOO when some dep index >= g: only then can the DFS post-order differ
// from 0..G-1 (otherwise every root's deps are already visited when the root loop reaches it).
// ---------------------------------------------------------------------------------------------------
// kCount (Kahn, K3): also counts the consumers of every producer (row lengths of the consumer CSR) while the dependencies are in
// registers - no second pass over dep[]; the forward-edge list (K5a seeds) is then not collected.
template <bool kCount>
__global__ void __launch_bounds__(kBlock) k_deps_t(const uint4* __restrict__ gates, uint32_t G, uint32_t node_bound, const uint32_t* __restrict__ prod1,
                                                   uint2* __restrict__ dep, uint32_t* __restrict__ seeds, uint32_t* __restrict__ cnt, uint32_t* __restrict__ scalars) {
  constexpr int kPend = 4;  // forward-edge gates a thread holds back: one warp-aggregated atomic per kPend iterations
  uint32_t pend[kPend];
  int np = 0;
  const int lane = threadIdx.x & 31;
  auto flush = [&]() {
    // warp-wide: offsets by an inclusive scan of the pending counts, one atomic for the warp's total
    if (!__ballot_sync(0xFFFFFFFFu, np != 0)) return;
    const uint32_t incl = warp_incl_scan((uint32_t)np, lane);
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    uint32_t base = 0;
    if (lane == 31) base = atomicAdd(scalars + S_SEEDN, total);
    base = __shfl_sync(0xFFFFFFFFu, base, 31) + incl - np;
#pragma unroll
    for (int i = 0; i < kPend; ++i) if (i < np) seeds[base + i] = pend[i];
    np = 0;
  };
  uint32_t f = 0;
  const uint32_t iters = (G + gridDim.x * kBlock - 1) / (gridDim.x * kBlock);  // full-warp iterations (warp-aggregated append below)
  for (uint32_t it = 0; it < iters; ++it) {
    const uint32_t g = it * gridDim.x * kBlock + blockIdx.x * kBlock + threadIdx.x;
    bool fwd = false;
    if (g < G) {
      uint4 gt = ldg_stream(gates + g);
      // ids >= node_bound were flagged by k_producer (F_BAD); stay memory-safe here, the call fails at the status read
      uint32_t d0 = gt.y < node_bound ? __ldg(prod1 + gt.y) - 1u : kNone;  // 0 -> kNone
      uint32_t d1 = gt.z < node_bound ? __ldg(prod1 + gt.z) - 1u : kNone;
      dep[g] = make_uint2(d0, d1);
      fwd = (d0 != kNone && d0 > g) || (d1 != kNone && d1 > g);
      if (fwd) f |= F_OOO;
      if (d0 == g || d1 == g) f |= F_SELF;
      if (kCount) {
        if (d0 != kNone) atomicAdd(cnt + d0, 1u);
        if (d1 != kNone && d1 != d0) atomicAdd(cnt + d1, 1u);  // lh == rh: one producer, listed once
      }
    }
    if (kCount) continue;
    // gates with a forward dependency are where the relaxation starts (K5a): collect them, the seed kernel does not
    // have to stream dep[] again to find them.  Held back per thread and appended once per kPend iterations: a shuffled gate
    // vector has a forward edge in every other gate, and one atomic per warp and iteration on the one counter then bounds the kernel.
    if (fwd) {
#pragma unroll
      for (int i = 0; i < kPend; ++i) if (i == np) pend[i] = g;
      ++np;
    }
    if ((it % kPend) == kPend - 1) flush();
  }
  if (!kCount) flush();
  uint32_t any_ooo = __syncthreads_or(f & F_OOO), any_self = __syncthreads_or(f & F_SELF);
  if (threadIdx.x == 0 && (any_ooo || any_self)) atomicOr(scalars + S_FLAGS, (any_ooo ? F_OOO : 0u) | (any_self ? F_SELF : 0u));
}

// generic get_deps form (topological_sort.rs:3-6): CSR rows with <= 2 entries
__global__ void __launch_bounds__(kBlock) k_deps_from_csr(const unsigned long long* __restrict__ dep_off, const uint32_t* __restrict__ dep_idx,
                                                          uint32_t n, uint2* __restrict__ dep, uint32_t* __restrict__ seeds, uint32_t* __restrict__ scalars) {
  uint32_t f = 0;
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < n; g += gridDim.x * kBlock) {
    unsigned long long a = dep_off[g], b = dep_off[g + 1];
    uint32_t d0 = kNone, d1 = kNone;
    if (b < a || b - a > 2) f |= F_BAD;
    else {
      if (b - a >= 1) d0 = dep_idx[a];
      if (b - a == 2) d1 = dep_idx[a + 1];
      if (d0 != kNone && d0 >= n) { f |= F_BAD; d0 = kNone; }
      if (d1 != kNone && d1 >= n) { f |= F_BAD; d1 = kNone; }
      if (b - a == 2 && d0 == kNone) f |= F_BAD;  // a literal 0xFFFFFFFF index is not representable
    }
    dep[g] = make_uint2(d0, d1);
    if ((d0 != kNone && d0 > g) || (d1 != kNone && d1 > g)) { f |= F_OOO; seeds[atomicAdd(scalars + S_SEEDN, 1u)] = g; }
    if (d0 == g || d1 == g) f |= F_SELF;
  }
  uint32_t any_ooo = __syncthreads_or(f & F_OOO), any_bad = __syncthreads_or(f & F_BAD), any_self = __syncthreads_or(f & F_SELF);
  if (threadIdx.x == 0 && (any_ooo || any_bad || any_self))
    atomicOr(scalars + S_FLAGS, (any_ooo ? F_OOO : 0u) | (any_bad ? F_BAD : 0u) | (any_self ? F_SELF : 0u));
}

__global__ void __launch_bounds__(kBlock) k_iota(uint32_t* __restrict__ a, uint32_t n) {
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) a[i] = i;
}
// order = 0..n-1 when no dependency points forward (topological_sort.rs:11-13: every root's deps are already visited)
__global__ void __launch_bounds__(kBlock) k_iota_if_identity(uint32_t* __restrict__ a, uint32_t n, const uint32_t* __restrict__ sc) {
  if (sort_wanted(sc) || (sc[S_FLAGS] & F_BAD)) return;
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) a[i] = i;
}
// K5 scratch: r = iota, size_off = 0, state = 0, inq = 0 (only when the sort has to run)
__global__ void __launch_bounds__(kBlock) k_sort_init(uint32_t n, uint32_t* __restrict__ r, uint32_t* __restrict__ size_off, uint8_t* __restrict__ state,
                                                      uint32_t* __restrict__ inq, const uint32_t* __restrict__ sc) {
  if (!sort_wanted(sc)) return;
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i <= n; i += gridDim.x * kBlock) {
    size_off[i] = 0;
    if (i < n) { r[i] = i; state[i] = 0; }
    if (i <= (n + 31) / 32) inq[i] = 0;
  }
}

// ---------------------------------------------------------------------------------------------------
// K5a: r[v] = index of the root of the `for i in 0..len` loop (topological_sort.rs:11-13) whose visit first
// reaches v = min(v, min over transitive consumers).  Monotone min-propagation along dependency edges,
// data-driven: whoever lowers r[x] is responsible for x's dependencies - it chases the lh chain itself and
// queues the rh side for the next round.  inq (bitmask) keeps a queue at <= one entry per item.
// r only ever decreases, so stale (too high) reads are harmless: they cost one redundant atomicMin.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void enqueue(uint32_t x, uint32_t* __restrict__ inq, uint32_t* __restrict__ q, uint32_t* __restrict__ qn) {
  uint32_t bit = 1u << (x & 31);
  __threadfence();  // the r[] update must be visible before the flag
  uint32_t old = atomicOr(inq + (x >> 5), bit);
  if (!(old & bit)) q[atomicAdd(qn, 1u)] = x;
}

__device__ __forceinline__ void relax_from(uint32_t cur, uint32_t val, const uint2* __restrict__ dep, uint32_t* __restrict__ r,
                                           uint32_t* __restrict__ inq, uint32_t* __restrict__ q, uint32_t* __restrict__ qn) {
  // invariant on entry: r[cur] <= val; cur's dependencies may still hold larger values
  while (cur != kNone) {
    uint2 d = dep[cur];
    uint32_t nxt = kNone;
    if (d.y != kNone && d.y != d.x && val < __ldcg(r + d.y)) {
      if (val < atomicMin(r + d.y, val)) enqueue(d.y, inq, q, qn);
    }
    if (d.x != kNone && val < __ldcg(r + d.x)) {
      if (val < atomicMin(r + d.x, val)) nxt = d.x;
    }
    cur = nxt;
  }
}

// Grid barrier of the cooperative kernels of the sort (all CTAs co-resident): bar.sync orders the CTA before thread 0's release,
// the acquire poll orders it - and through the second bar.sync the whole CTA - behind every other CTA's release.
__device__ __forceinline__ void sort_grid_bar(unsigned int* bar, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned int v;
    uint32_t spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (++spins > (1u << 26)) __trap();
    } while (v < epoch);
  }
  __syncthreads();
}

// relax_from with a bounded walk: after kHopCap hops the node reached is queued and continues in the next round, so one round costs
// O(queue x kHopCap) whatever the shape of the DAG (an unbounded walk down a 1 M-gate chain is a second of dependent loads).
// The seeds probe with a short cap (a reversed chain makes EVERY seed walk to the cap); the rounds use a long one, so that a walk
// inside a shuffled component (hundreds of hops) is not cut into several rounds - a round lasts as long as its longest walk.
constexpr uint32_t kSeedHopCap = 64, kRoundHopCap = 2048;
// The rh side of a lowered gate goes on a small per-thread stack and is walked by the same thread once the lh chain ends (the global
// queue takes the overflow): a label then crosses a whole shuffled component in ONE round instead of one rh hop per round
// (BASELINE config 5 shuffled: 101 rounds before).  The hop cap counts the whole walk; whatever is pending when it is reached is queued.
constexpr int kRelaxStack = 8;
__device__ __forceinline__ void relax_capped(uint32_t cur, uint32_t val, const uint2* dep, uint32_t* r, uint32_t* inq, uint32_t* q, uint32_t* qn, uint32_t* ncut,
                                             uint32_t cap) {
  uint32_t stk[kRelaxStack];
  int sp = 0;
  for (uint32_t hop = 0;; ++hop) {
    if (cur == kNone) {
      if (!sp) return;
      cur = stk[--sp];
    }
    if (hop == cap) {
      enqueue(cur, inq, q, qn);
      while (sp) enqueue(stk[--sp], inq, q, qn);
      atomicAdd(ncut, 1u);
      return;
    }
    uint2 d = __ldcg(dep + cur);
    uint32_t nxt = kNone;
    if (d.y != kNone && d.y != d.x && val < __ldcg(r + d.y)) {
      if (val < atomicMin(r + d.y, val)) {
        if (sp < kRelaxStack) stk[sp++] = d.y;
        else enqueue(d.y, inq, q, qn);
      }
    }
    if (d.x != kNone && val < __ldcg(r + d.x)) {
      if (val < atomicMin(r + d.x, val)) nxt = d.x;
    }
    cur = nxt;
  }
}

// K5a as ONE cooperative kernel: seeds, then data-driven rounds on device-resident queue counters with a grid barrier between
// rounds (a host-synchronised round costs ~50 us; a shuffled 10 M-gate vector needs ~90 of them).  When walks keep running into
// the hop cap - many of them after the second round, or any at all after kPatientRounds - the DAG is deep (a long forward chain) and
// label-by-label propagation is a latency chain: stage 2 computes, by pointer jumping in O(log depth) sweeps, the
// minimum of r[] along every gate's chain of smallest-index consumers (up[]: exact for forests, an upper bound otherwise - every
// value is still the index of a real transitive consumer), then the data-driven rounds finish from the violations that are left.
constexpr uint32_t kPatientRounds = 4, kMaxDataRounds = 2048;
__global__ void __launch_bounds__(kBlock) k_relax_loop(const uint2* dep, uint32_t n, const uint32_t* seeds, uint32_t* r, uint32_t* inq, uint32_t* q0, uint32_t* q1,
                                                       uint32_t* sc) {
  if (!sort_wanted(sc)) return;
  unsigned int epoch = 0;
  unsigned int* bar = reinterpret_cast<unsigned int*>(sc + S_BAR1);
  const uint32_t tid = blockIdx.x * kBlock + threadIdx.x, nth = gridDim.x * kBlock;
  // r[d] <= d always, so val = u can only lower r[d] along a forward edge (d > u): exactly the gates k_deps collected.
  // When every fourth gate holds a forward edge (a shuffled gate vector) the jumping stage runs FIRST: it replaces millions of
  // overlapping walks by O(log depth) sweeps, and the data-driven rounds only repair what the chains of smallest-index consumers miss.
  const uint32_t ns = __ldcg(sc + S_SEEDN);
  const bool dense_seeds = ns > n / 4;
  if (!dense_seeds) {
    for (uint32_t i = tid; i < ns; i += nth) { uint32_t u = __ldcg(seeds + i); relax_capped(u, u, dep, r, inq, q0, sc + S_QN0, sc + S_CAP0, kSeedHopCap); }
  }
  sort_grid_bar(bar, epoch);
  uint32_t slot = 0, rounds = 0;
  uint32_t* qin = q0;
  uint32_t* qout = q1;
  uint32_t capw = 0;  // the S_CAP word the previous phase counted into (the seeds: word 0)
  uint32_t prev_nq = 0xFFFFFFFFu;
  auto data_rounds = [&](bool patient) -> bool {  // true: drained; false (only when !patient): deep DAG, go pointer jumping
    for (uint32_t k = 0;; ++k) {
      const uint32_t nq = __ldcg(sc + S_QN0 + slot);
      if (!nq) return true;
      const uint32_t ncut = __ldcg(sc + S_CAP0 + capw);  // walks of the previous phase that ran into the hop cap
      // deep DAG?  most seeds walked into the (short) cap; many walks still run into the long one; some still do after a few
      // rounds (one label travelling down one long chain); or the queue stays long and does not shrink (one hop per round through
      // the rh operand: nothing is ever cut)
      const bool stuck = k >= 3 && nq > n / 64 && (unsigned long long)nq * 8 > (unsigned long long)prev_nq * 7;
      if (!patient && ((k == 0 && ncut > n / 8) || (k >= 1 && ncut > n / 64) || (k >= kPatientRounds && ncut) || stuck || k >= kMaxDataRounds)) return false;
      prev_nq = nq;
      const uint32_t nslot = (slot + 1) & 3, ncapw = (capw + 1) % 3;
      if (tid == 0) { sc[S_QN0 + ((nslot + 1) & 3)] = 0; sc[S_CAP0 + (ncapw + 1) % 3] = 0; }
      for (uint32_t i = tid; i < nq; i += nth) {
        uint32_t x = __ldcg(qin + i);
        atomicAnd(inq + (x >> 5), ~(1u << (x & 31)));
        __threadfence();  // clear the flag before sampling r[x]: a later lowering re-queues x
        relax_capped(x, __ldcg(r + x), dep, r, inq, qout, sc + S_QN0 + nslot, sc + S_CAP0 + ncapw, kRoundHopCap);
      }
      capw = ncapw;
      sort_grid_bar(bar, epoch);
      uint32_t* t = qin; qin = qout; qout = t;
      slot = nslot;
      ++rounds;
    }
  };
  if (dense_seeds || !data_rounds(false)) {
    // ---- stage 2: pointer jumping along the smallest-index consumer.  The queues are abandoned (violations are recollected below).
    uint32_t* upA = q0;
    uint32_t* upB = q1;
    sort_grid_bar(bar, epoch);  // every CTA has read the queue length that sent it here before the counters are reset
    for (uint32_t v = tid; v < n; v += nth) upA[v] = kNone;
    for (uint32_t w = tid; w < (n + 31) / 32 + 1; w += nth) inq[w] = 0;
    if (tid == 0) {
      sc[S_QN0] = sc[S_QN0 + 1] = sc[S_QN0 + 2] = sc[S_QN0 + 3] = 0;
      sc[S_CHG0] = sc[S_CHG0 + 1] = sc[S_CHG0 + 2] = 0;
      sc[S_CAP0] = sc[S_CAP0 + 1] = sc[S_CAP0 + 2] = 0;
      sc[S_ROBUST] = 1;
    }
    sort_grid_bar(bar, epoch);
    for (uint32_t u = tid; u < n; u += nth) {
      const uint2 d = __ldcg(dep + u);
      if (d.x != kNone && d.x != u) atomicMin(upA + d.x, u);
      if (d.y != kNone && d.y != d.x && d.y != u) atomicMin(upA + d.y, u);
    }
    sort_grid_bar(bar, epoch);
    for (uint32_t it = 0;; ++it) {
      // synchronous doubling on the pointers (upA read, upB written), r[] in place: a fresher r[a] only covers MORE of a's chain
      uint32_t chg = 0;
      for (uint32_t v = tid; v < n; v += nth) {
        const uint32_t a = __ldcg(upA + v);
        uint32_t nu = kNone;
        if (a != kNone) {
          const uint32_t ra = __ldcg(r + a);
          if (ra < __ldcg(r + v)) atomicMin(r + v, ra);
          nu = __ldcg(upA + a);
          chg = 1;
        }
        upB[v] = nu;
      }
      if (__syncthreads_or(chg) && threadIdx.x == 0) atomicOr(sc + S_CHG0 + it % 3, 1u);
      if (tid == 0) sc[S_CHG0 + (it + 1) % 3] = 0;
      sort_grid_bar(bar, epoch);
      uint32_t* t = upA; upA = upB; upB = t;
      ++rounds;
      if (!__ldcg(sc + S_CHG0 + it % 3)) break;
      if (it > 64) break;  // 2^64 > any depth: cannot happen
    }
    // the chain of smallest-index consumers misses the other consumers of a shared gate: collect what is still violated ...
    qin = q0; qout = q1; slot = 0; capw = 0;   // (up[] is dead: the arrays are queues again)
    for (uint32_t u = tid; u < n; u += nth) {
      const uint2 d = __ldcg(dep + u);
      const uint32_t ru = __ldcg(r + u);
      if (d.x != kNone && ru < __ldcg(r + d.x)) { if (ru < atomicMin(r + d.x, ru)) enqueue(d.x, inq, qin, sc + S_QN0); }
      if (d.y != kNone && d.y != d.x && ru < __ldcg(r + d.y)) { if (ru < atomicMin(r + d.y, ru)) enqueue(d.y, inq, qin, sc + S_QN0); }
    }
    sort_grid_bar(bar, epoch);
    // ... and finish data-driven (every round strictly lowers some r[]: it terminates)
    data_rounds(true);
  }
  if (tid == 0) sc[S_RELAX_ROUNDS] = rounds;
}

// K5b: block sizes.  size[root] = number of items first reached from root.
// Runs of equal roots inside a warp (neighbouring gates usually belong to one block; a single giant tree is ONE run) are added with
// one atomic: 1 M same-address atomics of a one-tree circuit took 0.65 ms.
__global__ void __launch_bounds__(kBlock) k_sizes(const uint32_t* __restrict__ r, uint32_t n, uint32_t* __restrict__ size, const uint32_t* __restrict__ sc) {
  if (!sort_wanted(sc) || sort_parked(sc)) return;
  const int lane = threadIdx.x & 31;
  const uint32_t iters = (n + gridDim.x * kBlock - 1) / (gridDim.x * kBlock);
  for (uint32_t it = 0; it < iters; ++it) {
    const uint32_t v = it * gridDim.x * kBlock + blockIdx.x * kBlock + threadIdx.x;
    const bool valid = v < n;
    const uint32_t rv = valid ? r[v] : kNone;
    const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, rv, 1);
    const bool head = valid && (lane == 0 || rv != prev);
    const uint32_t heads = __ballot_sync(0xFFFFFFFFu, head), vmask = __ballot_sync(0xFFFFFFFFu, valid);
    if (head) {
      const uint32_t after = heads & ~((2u << lane) - 1u);          // run heads behind this one
      const int end = after ? __ffs(after) - 1 : __popc(vmask);     // (valid lanes are a prefix of the warp)
      atomicAdd(size + rv, (uint32_t)(end - lane));
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Single-pass exclusive scan (decoupled look-back).  Tile index = blockIdx.x: CTAs are dispatched in index order and run to
// completion, so a tile only ever waits on tiles that are already resident (the scheme of CUB's DeviceScan).  A global ticket
// per tile was measured instead: same-address atomics with a return value retire at one per ~18 ns, which alone bounded
// every look-back kernel here (2 446 / 9 784 / 40 905 tiles -> 47 / 197 / 672 us).
// tile_state word = (status << 32) | value, status 0 = empty, 1 = tile aggregate, 2 = inclusive prefix.
// ---------------------------------------------------------------------------------------------------
constexpr unsigned long long kStAgg = 1ull << 32, kStInc = 2ull << 32;


// scan_tile_prefix returns the exclusive prefix of this thread's `thread_sum` over the whole grid-wide sequence of tiles.
// Must be called by all kBlock threads.  s_mem: >= 10 u32 of shared memory.
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

// Decoupled look-back for tile > 0, called by one full warp: sum of the values of all tiles before `tile`.
// State word = status << kShift | value  (status 0 = empty, 1 = tile aggregate, 2 = inclusive prefix).
// Every lane keeps kLookWin states in flight (window = 32 * kLookWin tiles per round trip).  Measured with
// tools/lookback_bench.cu on B200 (10 M u32): the look-back adds ~15 us + ~5 ns per tile on top of the 15-18 us the data
// movement takes, and wider windows are slower (more polling traffic), so kLookWin = 1 and the tiles are made large instead.
constexpr int kLookWin = 1;
template <int kShift>
__device__ __forceinline__ unsigned long long lookback_exclusive(const unsigned long long* __restrict__ tile_state, uint32_t tile, int lane) {
  constexpr unsigned long long kValMask = (1ull << kShift) - 1;
  unsigned long long prefix = 0;
  long long p = (long long)tile - 1;
  uint32_t spins = 0;
  while (true) {
    unsigned long long v[kLookWin];
#pragma unroll
    for (int j = 0; j < kLookWin; ++j) {
      long long idx = p - lane - 32 * j;
      v[j] = idx >= 0 ? ld_volatile_u64(tile_state + idx) : (2ull << kShift);  // before tile 0: an inclusive prefix of 0
    }
#pragma unroll
    for (int j = 0; j < kLookWin; ++j) {
      long long idx = p - lane - 32 * j;
      while (true) {
        uint32_t st = (uint32_t)(v[j] >> kShift);
        uint32_t empty = __ballot_sync(0xFFFFFFFFu, st == 0), inc = __ballot_sync(0xFFFFFFFFu, st == 2);
        unsigned long long val = v[j] & kValMask;
        if (inc) {
          int first = __ffs(inc) - 1;  // nearest predecessor holding an inclusive prefix
          if (!(empty & ((1u << first) - 1u))) return prefix + warp_sum_u64(lane <= first ? val : 0ull);
        } else if (!empty) {
          prefix += warp_sum_u64(val);
          break;  // next 32 predecessors
        }
        if (st == 0) v[j] = ld_volatile_u64(tile_state + idx);
        if (++spins > (1u << 25)) __trap();  // tens of seconds of polling: a predecessor never published (fail loudly, do not hang)
      }
    }
    p -= 32 * kLookWin;
  }
}

__device__ __forceinline__ uint32_t scan_tile_prefix(uint32_t tile, uint32_t thread_sum, unsigned long long* __restrict__ tile_state,
                                                     uint32_t* s_mem, uint32_t* total_out, bool is_last_tile) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = warp_incl_scan(thread_sum, lane);
  if (lane == 31) s_mem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < (kBlock / 32) ? s_mem[lane] : 0u;
    uint32_t wi = warp_incl_scan(w, lane);
    uint32_t block_agg = __shfl_sync(0xFFFFFFFFu, wi, (kBlock / 32) - 1);
    if (lane < (kBlock / 32)) s_mem[lane] = wi - w;  // exclusive warp offsets
    uint32_t prefix = 0;
    if (tile == 0) {
      if (lane == 0) st_volatile_u64(tile_state, kStInc | block_agg);
    } else {
      if (lane == 0) st_volatile_u64(tile_state + tile, kStAgg | block_agg);
      prefix = (uint32_t)lookback_exclusive<32>(tile_state, tile, lane);
      if (lane == 0) st_volatile_u64(tile_state + tile, kStInc | (unsigned long long)(prefix + block_agg));
    }
    if (lane == 0) {
      s_mem[8] = prefix;
      if (is_last_tile && total_out) *total_out = prefix + block_agg;
    }
  }
  __syncthreads();
  return s_mem[8] + s_mem[warp] + (incl - thread_sum);
}

// exclusive scan of src[0..n) into dst[0..n) (may alias); the total goes to *total_out (default dst[n]).  128-bit accesses.
// guard_kind: 0 none, 1 = part of the sort (guard = its scalars), 2 = part of the wire numbering.
constexpr int kScanItems = 16;  // 4096 elements per tile: fewer links in the look-back chain (8192 was measured: 49 us against 52 us for the two build scans - not worth the registers)
template <bool kPopc>  // kPopc: scan popcount(src[i]) instead of src[i] (rank structure over a bitmap)
__device__ __forceinline__ void scan_u32_body(const uint32_t* src, uint32_t* dst, uint32_t n, unsigned long long* __restrict__ tile_state,
                                              uint32_t* __restrict__ total_out, uint32_t* s_mem) {
  const uint32_t tiles = (n + kBlock * kScanItems - 1) / (kBlock * kScanItems);
  const uint32_t tile = blockIdx.x;
  uint32_t base = tile * (kBlock * kScanItems) + threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  if (base + kScanItems <= n) {
#pragma unroll
    for (int i = 0; i < kScanItems; i += 4) {
      uint4 x = *reinterpret_cast<const uint4*>(src + base + i);
      v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) v[i] = base + i < n ? src[base + i] : 0u;
  }
  if (kPopc) {
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) v[i] = __popc(v[i]);
  }
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) sum += v[i];
  uint32_t ex = scan_tile_prefix(tile, sum, tile_state, s_mem, total_out ? total_out : dst + n, tile == tiles - 1);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) { uint32_t t = v[i]; v[i] = ex; ex += t; }
  if (base + kScanItems <= n) {
#pragma unroll
    for (int i = 0; i < kScanItems; i += 4) *reinterpret_cast<uint4*>(dst + base + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) if (base + i < n) dst[base + i] = v[i];
  }
}
template <bool kPopc>
__global__ void __launch_bounds__(kBlock) k_scan_u32_t(const uint32_t* src, uint32_t* dst, uint32_t n, unsigned long long* __restrict__ tile_state,
                                                       uint32_t* __restrict__ total_out, const uint32_t* __restrict__ guard, int guard_kind) {
  __shared__ uint32_t s_mem[10];
  if (guard_kind == 1 && (!sort_wanted(guard) || sort_parked(guard))) return;
  if (guard_kind == 2 && tail_parked(guard)) return;
  scan_u32_body<kPopc>(src, dst, n, tile_state, total_out, s_mem);
}
// up to three independent in-place scans of equal length in ONE launch (blockIdx.y picks the array): the emitter's per-tile gate /
// connection / implicit-operand counts - three back-to-back launches of a latency-bound 10-CTA kernel otherwise
struct ScanJobs { uint32_t* a[3]; unsigned long long* state[3]; };
__global__ void __launch_bounds__(kBlock) k_scan_u32_multi(ScanJobs jobs, uint32_t n) {
  __shared__ uint32_t s_mem[10];
  const int y = blockIdx.y;  // (selects instead of a dynamic index: the parameter struct stays in constant memory)
  uint32_t* a = y == 0 ? jobs.a[0] : y == 1 ? jobs.a[1] : jobs.a[2];
  unsigned long long* st = y == 0 ? jobs.state[0] : y == 1 ? jobs.state[1] : jobs.state[2];
  scan_u32_body<false>(a, a, n, st, nullptr, s_mem);
}

// ---------------------------------------------------------------------------------------------------
// K5c: emit.  Blocks of one item are written straight away; larger blocks go to the heavy list and are
// walked by k_tree_dfs, one thread per block, reproducing topological_sort_visit (topological_sort.rs:23-50)
// restricted to the block: an item with r < root was emitted by an earlier root ("visited"), r == root is
// tracked in state[].  The DFS stack lives in the block's own slice of order[] (it grows down from the end
// while emitted items grow up from the start; |stack| + |emitted| <= block size).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_roots(const uint32_t* __restrict__ r, const uint32_t* __restrict__ off, uint32_t n,
                                                  const uint2* __restrict__ dep, uint32_t* __restrict__ order,
                                                  uint32_t* __restrict__ heavy, uint32_t* __restrict__ scalars) {
  if (!sort_wanted(scalars) || sort_parked(scalars)) return;
  const bool check_self = scalars[S_FLAGS] & F_SELF;
  for (uint32_t v = blockIdx.x * kBlock + threadIdx.x; v < n; v += gridDim.x * kBlock) {
    if (r[v] != v) continue;
    uint32_t o = off[v], sz = off[v + 1] - o;
    if (sz == 1) {
      if (check_self) {
        uint2 d = dep[v];
        if (d.x == v || d.y == v)  // visit(v) -> visit(v) while visiting[v]  => "detected at i=v"
          atomicMin(reinterpret_cast<unsigned long long*>(scalars + S_ERR_LO), ((unsigned long long)v << 32) | v);
      }
      order[o] = v;
    } else {
      heavy[atomicAdd(scalars + S_HEAVYN, 1u)] = v;
      if (sz >= kBigBlock) atomicAdd(scalars + S_BIGBLK, 1u);  // k_tree_blocks looks at it: a one-thread DFS would take sz x ~6 us
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// K5c for BIG blocks that are trees (every gate of the block has exactly one consumer inside the block: a reversed chain, a
// forward chain through either operand, any fan-in tree).  There the DFS post-order is a closed formula,
//     post(v) = off[R] + L(v) + size(v) - 1,      L(v) = sum over the ancestors a of v (v included) that are rh children of
//                                                        size(lh sibling subtree)   (topological_sort.rs:42-47: lh before rh),
// and both size() and L() are pointer-jumping computations: O(log depth) sweeps instead of size x ~6 us of dependent loads in one
// thread (a 1 M-gate chain: 6 s; the reference's recursion: 0.15 s).  One cooperative kernel; it exits at once when k_roots saw no
// big block.  Blocks with a shared gate (in-block in-degree 2: a DAG, where "who visits first" is the sequential part of the
// problem) and blocks holding a cycle are flagged in sz[R] and left to k_tree_dfs.
//   indeg, list: the two relaxation queues (dead by now);  par / anc0 / anc1 / v0 / v1 / sz: six more gate-indexed arrays.
// ---------------------------------------------------------------------------------------------------
struct TreeArrays { uint32_t *indeg, *list, *par, *anc0, *anc1, *v0, *v1, *sz; };
constexpr uint32_t kNotATree = 0xFFFFFFFFu;
__global__ void __launch_bounds__(kBlock) k_tree_blocks(const uint2* dep, const uint32_t* r, const uint32_t* off, uint32_t n, uint32_t* order, TreeArrays t,
                                                        uint32_t* sc) {
  if (!sort_wanted(sc) || __ldcg(sc + S_BIGBLK) == 0) return;
  unsigned int epoch = 0;
  unsigned int* bar = reinterpret_cast<unsigned int*>(sc + S_BAR2);
  const uint32_t tid = blockIdx.x * kBlock + threadIdx.x, nth = gridDim.x * kBlock;
  const int lane = threadIdx.x & 31;
  auto big = [&](uint32_t R) { return __ldcg(off + R + 1) - __ldcg(off + R) >= kBigBlock; };
  for (uint32_t v = tid; v < n; v += nth) { t.indeg[v] = 0; t.sz[v] = 0; }
  if (tid == 0) { sc[S_CHG0] = sc[S_CHG0 + 1] = sc[S_CHG0 + 2] = 0; sc[S_BIGN] = 0; }
  sort_grid_bar(bar, epoch);
  // T1: parents and in-block in-degrees of the gates of big blocks
  for (uint32_t u = tid; u < n; u += nth) {
    const uint32_t R = __ldcg(r + u);
    if (!big(R)) continue;
    const uint2 d = __ldcg(dep + u);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t x = j ? d.y : d.x;
      if (x == kNone || (j && d.y == d.x) || __ldcg(r + x) != R) continue;
      t.par[x] = u | (j ? 0x80000000u : 0u);  // (with in-degree 1 there is one writer)
      if (atomicAdd(t.indeg + x, 1u) >= 1u || x == R) t.sz[R] = kNotATree;  // shared gate / a cycle through the root: k_tree_dfs
    }
  }
  sort_grid_bar(bar, epoch);
  // T2: the gates of big TREE blocks: list, initial ancestors and counts
  {
    const uint32_t iters = (n + nth - 1) / nth;
    for (uint32_t it = 0; it < iters; ++it) {
      const uint32_t v = it * nth + tid;
      bool in = false;
      if (v < n) {
        const uint32_t R = __ldcg(r + v);
        in = big(R) && __ldcg(t.sz + R) != kNotATree;
        if (in) { t.anc0[v] = v == R ? kNone : (__ldcg(t.par + v) & 0x7FFFFFFFu); t.v0[v] = 1; }
      }
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, in);
      if (m) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(sc + S_BIGN, (uint32_t)__popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (in) t.list[base + __popc(m & ((1u << lane) - 1))] = v;
      }
    }
  }
  sort_grid_bar(bar, epoch);
  const uint32_t nb = __ldcg(sc + S_BIGN);
  if (nb == 0) return;  // every big block is a DAG: nothing to do here (uniform)
  uint32_t *a0 = t.anc0, *a1 = t.anc1, *c0 = t.v0, *c1 = t.v1;
  // T3: subtree sizes.  cnt_k[v] = gates of v's subtree at distance < 2^k, anc_k[v] = the ancestor at distance 2^k:
  //     cnt_{k+1}[a] = cnt_k[a] + sum of cnt_k[v] over the v with anc_k[v] = a;  anc_{k+1}[v] = anc_k[anc_k[v]]
  for (uint32_t it = 0;; ++it) {
    for (uint32_t i = tid; i < nb; i += nth) { const uint32_t v = __ldcg(t.list + i); c1[v] = __ldcg(c0 + v); }
    if (tid == 0) sc[S_CHG0 + (it + 1) % 3] = 0;
    sort_grid_bar(bar, epoch);
    uint32_t chg = 0;
    for (uint32_t i = tid; i < nb; i += nth) {
      const uint32_t v = __ldcg(t.list + i), a = __ldcg(a0 + v);
      uint32_t na = kNone;
      if (a != kNone) { atomicAdd(c1 + a, __ldcg(c0 + v)); na = __ldcg(a0 + a); chg = 1; }
      a1[v] = na;
    }
    if (__syncthreads_or(chg) && threadIdx.x == 0) atomicOr(sc + S_CHG0 + it % 3, 1u);
    sort_grid_bar(bar, epoch);
    uint32_t* x = a0; a0 = a1; a1 = x;
    x = c0; c0 = c1; c1 = x;
    if (!__ldcg(sc + S_CHG0 + it % 3) || it > 40) break;
  }
  // sizes are final in c0; keep them in sz[], start the offset sums: w(v) = size(lh sibling) for an rh child whose lh sibling is a
  // child too, else 0
  for (uint32_t i = tid; i < nb; i += nth) t.sz[__ldcg(t.list + i)] = __ldcg(c0 + __ldcg(t.list + i));
  sort_grid_bar(bar, epoch);
  if (tid == 0) sc[S_CHG0] = sc[S_CHG0 + 1] = sc[S_CHG0 + 2] = 0;  // (behind a barrier: every CTA has left the loop above)
  for (uint32_t i = tid; i < nb; i += nth) {
    const uint32_t v = __ldcg(t.list + i), R = __ldcg(r + v);
    uint32_t w = 0, a = kNone;
    if (v != R) {
      const uint32_t pw = __ldcg(t.par + v);
      a = pw & 0x7FFFFFFFu;
      if (pw & 0x80000000u) {  // rh child: everything under the lh sibling is emitted first
        const uint32_t c1n = __ldcg(dep + a).x;
        if (c1n != kNone && c1n != v && __ldcg(r + c1n) == R) w = __ldcg(t.sz + c1n);
      }
    }
    c0[v] = w;
    a0[v] = a;
  }
  sort_grid_bar(bar, epoch);
  // T4: L(v) = w(v) + L(parent): pull-doubling (both arrays ping-pong)
  for (uint32_t it = 0;; ++it) {
    uint32_t chg = 0;
    for (uint32_t i = tid; i < nb; i += nth) {
      const uint32_t v = __ldcg(t.list + i), a = __ldcg(a0 + v);
      uint32_t L = __ldcg(c0 + v), na = kNone;
      if (a != kNone) { L += __ldcg(c0 + a); na = __ldcg(a0 + a); chg = 1; }
      c1[v] = L;
      a1[v] = na;
    }
    if (__syncthreads_or(chg) && threadIdx.x == 0) atomicOr(sc + S_CHG0 + it % 3, 1u);
    if (tid == 0) sc[S_CHG0 + (it + 1) % 3] = 0;
    sort_grid_bar(bar, epoch);
    uint32_t* x = a0; a0 = a1; a1 = x;
    x = c0; c0 = c1; c1 = x;
    if (!__ldcg(sc + S_CHG0 + it % 3) || it > 40) break;
  }
  // T5: sorted.push order (topological_sort.rs:47)
  for (uint32_t i = tid; i < nb; i += nth) {
    const uint32_t v = __ldcg(t.list + i), R = __ldcg(r + v);
    order[__ldcg(off + R) + __ldcg(c0 + v) + __ldcg(t.sz + v) - 1u] = v;
  }
}

__global__ void __launch_bounds__(128) k_tree_dfs(const uint32_t* __restrict__ heavy,
                                                  const uint2* __restrict__ dep, const uint32_t* __restrict__ r,
                                                  const uint32_t* __restrict__ off, uint8_t* __restrict__ state,
                                                  uint32_t* __restrict__ order, const uint32_t* __restrict__ tree_sz, uint32_t* scalars) {
  if (!sort_wanted(scalars) || sort_parked(scalars)) return;
  uint32_t nh = scalars[S_HEAVYN];
  const bool trees_done = scalars[S_BIGBLK] != 0;  // k_tree_blocks ran: it emitted the big blocks it did not flag
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nh; i += gridDim.x * blockDim.x) {
    const uint32_t R = heavy[i];
    const uint32_t base = off[R], end = off[R + 1];
    if (trees_done && end - base >= kBigBlock && tree_sz[R] != kNotATree) continue;
    uint32_t emit = base, top = end;
    // state: 0 unvisited, 1 entered (lh next), 2 (rh next), 3 (both examined, emit next), 4 visited.
    // The gate being visited lives in REGISTERS (v, its progress s, its dependency pair dd); the block's slice of order[] holds the
    // emitted gates (growing up) and the stack of its ancestors (growing down).  A step is then one round trip for the probe of a
    // dependency (r[], state[] and - speculatively - its own dependency pair, three independent loads) and two for a return to the
    // parent, instead of five dependent loads per step and three steps per gate.
    uint32_t v = R, s = 1, pv = kNone;
    uint2 dd = dep[R];
    state[R] = 1;
    bool popping = false;  // the parent's index has been read from the stack, its state / dependency pair are still to be fetched
    // The 32 lanes of a warp walk 32 different blocks and are rarely in the same kind of step.  Every iteration therefore has ONE place
    // where loads are issued - a dependency probe, the stack read of a return, or the parent's record - and the results are consumed
    // afterwards: the lanes' round trips overlap instead of following one another branch by branch (that serialisation was ~2/3 of
    // this kernel's time on the shuffled BASELINE vector).
    while (true) {
      uint32_t d = kNone;
      bool want_probe = false, want_pop1 = false;
      const bool want_pop2 = popping;
      if (!popping) {
        if (s <= 2) {
          d = (s == 1) ? dd.x : dd.y;
          ++s;
          want_probe = d != kNone;
        } else {
          order[emit++] = v;  // sorted.push(i)
          state[v] = 4;       // visited[i] = true
          if (top == end) break;
          want_pop1 = true;
        }
      }
      // ---- issue
      uint32_t rd = 0, nv = 0;
      uint8_t sd = 0, ps = 0;
      uint2 dn = make_uint2(kNone, kNone), pdd = make_uint2(kNone, kNone);
      if (want_probe) { rd = r[d]; sd = state[d]; dn = dep[d]; }  // (dn is used only when the walk descends into d)
      if (want_pop1) nv = order[top];
      if (want_pop2) { ps = state[pv]; pdd = dep[pv]; }
      // ---- consume
      if (want_probe && rd == R) {  // (rd != R: emitted by an earlier root - visited)
        if (sd == 0) {              // descend: the current gate becomes an ancestor
          state[v] = (uint8_t)s;
          order[--top] = v;
          v = d; s = 1; dd = dn;
          state[d] = 1;
        } else if (sd < 4) {        // visiting[d]  (topological_sort.rs:34-38)
          atomicMin(reinterpret_cast<unsigned long long*>(scalars + S_ERR_LO), ((unsigned long long)R << 32) | d);
          break;
        }
      }
      if (want_pop1) { pv = nv; ++top; popping = true; }
      else if (want_pop2) { v = pv; s = ps; dd = pdd; popping = false; }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// K6: wire numbering (compiler.rs:388-449)
// ---------------------------------------------------------------------------------------------------
// I/O node lists (compiler.rs:392-395, 446-449): `insert(node, next_wire_id++)` in list order, a node listed twice
// keeps the LAST id.  On the device: zero the listed words, then RED.MAX(base + i).
__global__ void __launch_bounds__(kBlock) k_io_set(const uint32_t* __restrict__ nodes, uint32_t n, uint32_t node_bound, uint32_t value,
                                                   uint32_t* __restrict__ wire, uint32_t* __restrict__ scalars) {
  bool bad = false;
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    uint32_t nd = nodes[i];
    if (nd >= node_bound) bad = true;
    else wire[nd] = value;
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(scalars + S_FLAGS, (uint32_t)F_BAD_IO);
}
__global__ void __launch_bounds__(kBlock) k_io_max(const uint32_t* __restrict__ nodes, uint32_t n, uint32_t node_bound, uint32_t base,
                                                   const uint32_t* __restrict__ base_dev, uint32_t* __restrict__ wire) {
  uint32_t b = base + (base_dev ? *base_dev : 0u);
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    uint32_t nd = nodes[i];
    if (nd < node_bound) atomicMax(wire + nd, b + i);
  }
}

// pass 1: earliest appearance p = 3*pos+slot of every not-yet-numbered node, over the SORTED gate stream.
// RED.MIN per slot: inputs / pending outputs hold smaller words and are left untouched.
// Every thread keeps kWireIlp independent gates in flight: the order[] loads, then the gate loads, then the wire[] probes are
// issued as batches (a single gate is a chain of three dependent memory round trips).
// (Skipping the operand probes of nodes that have a producer - their first appearance is the producer's out slot - by reading dep[]
//  instead was measured: 0.085 ms against 0.080 ms, the 8-byte dep[] stream costs what the two L2-resident probes cost.)
constexpr int kWireIlp = 4;
__global__ void __launch_bounds__(kBlock) k_wire_first(const uint4* __restrict__ gates, const uint32_t* __restrict__ order_arr, uint32_t G,
                                                       uint32_t* __restrict__ wire, const uint32_t* __restrict__ sc) {
  if (tail_parked(sc)) return;
  const uint32_t* __restrict__ order = sort_wanted(sc) ? order_arr : nullptr;
  const uint32_t stride = gridDim.x * kBlock;
  for (uint32_t k0 = blockIdx.x * kBlock + threadIdx.x; k0 < G; k0 += stride * kWireIlp) {
    uint32_t g[kWireIlp];
    uint4 gt[kWireIlp];
    uint32_t w[kWireIlp][3];
#pragma unroll
    for (int i = 0; i < kWireIlp; ++i) {
      uint32_t k = min(k0 + i * stride, G - 1);  // clamped duplicates are harmless: RED.MIN of the same value
      g[i] = order ? __ldg(order + k) : k;
    }
#pragma unroll
    for (int i = 0; i < kWireIlp; ++i) gt[i] = order ? __ldg(gates + g[i]) : ldg_stream(gates + g[i]);
    // Read before the RED: words only ever decrease, so a (possibly stale) value <= p proves the RED is a no-op.
    // This removes the serialisation on hot nodes (a shared input such as a key feeds millions of gates), and a RED on a sector
    // that is not in L2 yet is 3.5x slower than load-then-RED (measured: 0.28 ms instead of 0.08 ms for this kernel).
#pragma unroll
    for (int i = 0; i < kWireIlp; ++i) { w[i][0] = wire[gt[i].y]; w[i][1] = wire[gt[i].z]; w[i][2] = wire[gt[i].w]; }
#pragma unroll
    for (int i = 0; i < kWireIlp; ++i) {
      uint32_t k = min(k0 + i * stride, G - 1);
      uint32_t p = kFirstTag | (3u * k);
      if (w[i][0] > p) atomicMin(wire + gt[i].y, p);
      if (gt[i].z != gt[i].y && w[i][1] > p + 1) atomicMin(wire + gt[i].z, p + 1);
      if (gt[i].w != gt[i].y && gt[i].w != gt[i].z && w[i][2] > p + 2) atomicMin(wire + gt[i].w, p + 2);
    }
  }
}

// pass 2 (compiler.rs:440-441): a node that still carries a tag was first seen at position p = 3*pos+slot and gets the wire
// n_in + #{first appearances before p}.  Done from the NODE side, fully coalesced over wire[]:
//   k_wire_mark    bitmap[p] = 1 for every tagged node              (3G bits: L2-resident)
//   k_scan_u32<popc>  pre[w] = number of set bits before word w;  total = number of intermediate wires
//   k_wire_assign  wire[node] = n_in + pre[p / 32] + popc(bitmap[p / 32] below bit p % 32)
// (A gate-side single-pass scan - three wire[] gathers per gate chained to a decoupled look-back - took 0.165-0.2 ms at 10 M
//  gates: its tiles spend half their life waiting for the prefix.)
__global__ void __launch_bounds__(kBlock) k_wire_mark(const uint32_t* __restrict__ wire, uint32_t node_bound, uint32_t* __restrict__ bitmap,
                                                      const uint32_t* __restrict__ sc) {
  if (tail_parked(sc)) return;
  const uint32_t n4 = (reinterpret_cast<uintptr_t>(wire) & 15) ? 0u : node_bound / 4;  // 128-bit path needs a 16-byte aligned map
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n4; i += gridDim.x * kBlock) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(wire) + i);
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((w[j] & kFirstTag) && w[j] != kNone) { uint32_t p = w[j] & ~kFirstTag; atomicOr(bitmap + (p >> 5), 1u << (p & 31)); }
  }
  for (uint32_t nd = n4 * 4 + blockIdx.x * kBlock + threadIdx.x; nd < node_bound; nd += gridDim.x * kBlock) {
    uint32_t w = wire[nd];
    if ((w & kFirstTag) && w != kNone) { uint32_t p = w & ~kFirstTag; atomicOr(bitmap + (p >> 5), 1u << (p & 31)); }
  }
}
__device__ __forceinline__ uint32_t wire_rank(uint32_t w, uint32_t n_in, const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ pre) {
  if (!(w & kFirstTag) || w == kNone) return w;  // input / output / never seen
  uint32_t p = w & ~kFirstTag;
  return n_in + __ldg(pre + (p >> 5)) + __popc(__ldg(bitmap + (p >> 5)) & ((1u << (p & 31)) - 1u));
}
__global__ void __launch_bounds__(kBlock) k_wire_assign(uint32_t* __restrict__ wire, uint32_t node_bound, uint32_t n_in,
                                                        const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ pre,
                                                        const uint32_t* __restrict__ sc) {
  if (tail_parked(sc)) return;
  const uint32_t n4 = (reinterpret_cast<uintptr_t>(wire) & 15) ? 0u : node_bound / 4;
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n4; i += gridDim.x * kBlock) {
    uint4 v = reinterpret_cast<const uint4*>(wire)[i];
    v.x = wire_rank(v.x, n_in, bitmap, pre);
    v.y = wire_rank(v.y, n_in, bitmap, pre);
    v.z = wire_rank(v.z, n_in, bitmap, pre);
    v.w = wire_rank(v.w, n_in, bitmap, pre);
    reinterpret_cast<uint4*>(wire)[i] = v;
  }
  for (uint32_t nd = n4 * 4 + blockIdx.x * kBlock + threadIdx.x; nd < node_bound; nd += gridDim.x * kBlock)
    wire[nd] = wire_rank(wire[nd], n_in, bitmap, pre);
}

// K7: gather (compiler.rs:452-464).  op stays numeric; the host maps it to the strum Display token.
constexpr int kGatherIlp = 2;
// (sorted positions [base, base + cnt): the whole circuit, or one chunk of it when the copy-out of the previous chunk runs beside it)
__global__ void __launch_bounds__(kBlock) k_gather(const uint4* __restrict__ gates, const uint32_t* __restrict__ order_arr, uint32_t base, uint32_t cnt,
                                                   const uint32_t* __restrict__ wire, uint4* __restrict__ new_gates, const uint32_t* __restrict__ sc) {
  if (tail_parked(sc)) return;
  const uint32_t* __restrict__ order = sort_wanted(sc) ? order_arr : nullptr;
  const uint32_t stride = gridDim.x * kBlock;
  const uint32_t G = base + cnt;
  for (uint32_t k0 = base + blockIdx.x * kBlock + threadIdx.x; k0 < G; k0 += stride * kGatherIlp) {
    uint32_t g[kGatherIlp];
    uint4 gt[kGatherIlp];
#pragma unroll
    for (int i = 0; i < kGatherIlp; ++i) { uint32_t k = min(k0 + i * stride, G - 1); g[i] = order ? __ldg(order + k) : k; }
#pragma unroll
    for (int i = 0; i < kGatherIlp; ++i) gt[i] = order ? __ldg(gates + g[i]) : ldg_stream(gates + g[i]);
#pragma unroll
    for (int i = 0; i < kGatherIlp; ++i) { gt[i].y = __ldg(wire + gt[i].y); gt[i].z = __ldg(wire + gt[i].z); gt[i].w = __ldg(wire + gt[i].w); }
#pragma unroll
    for (int i = 0; i < kGatherIlp; ++i) { uint32_t k = k0 + i * stride; if (k < G) stg_stream(new_gates + k, gt[i]); }
  }
}

// K7 for a sharded build: the gather applies the global offsets derived from the all-gathered counts (rank-major
// {n_in, n_mid, n_out, G}) and shifts order[] to global gate indices in the same pass.  order == nullptr: identity order.
__global__ void __launch_bounds__(kBlock) k_gather_global(const uint4* __restrict__ gates, uint32_t* __restrict__ order, uint32_t identity, uint32_t G,
                                                          const uint32_t* __restrict__ wire, uint4* __restrict__ new_gates,
                                                          const unsigned long long* __restrict__ counts, uint32_t rank, uint32_t world) {
  uint32_t n_in = 0xFFFFFFFFu, n_mid = 0, off_in = 0, off_mid = 0, off_out = 0, gate_base = 0;
  if (counts) {
    unsigned long long tot_in = 0, tot_mid = 0, pre_in = 0, pre_mid = 0, pre_out = 0, pre_g = 0;
    for (uint32_t r = 0; r < world; ++r) {
      tot_in += counts[4 * r];
      tot_mid += counts[4 * r + 1];
      if (r < rank) { pre_in += counts[4 * r]; pre_mid += counts[4 * r + 1]; pre_out += counts[4 * r + 2]; pre_g += counts[4 * r + 3]; }
    }
    n_in = (uint32_t)counts[4 * rank];
    n_mid = (uint32_t)counts[4 * rank + 1];
    off_in = (uint32_t)pre_in;
    off_mid = (uint32_t)(tot_in + pre_mid) - n_in;
    off_out = (uint32_t)(tot_in + tot_mid + pre_out) - n_in - n_mid;
    gate_base = (uint32_t)pre_g;
  }
  auto fix = [&](uint32_t w) { return !counts ? w : (w < n_in ? w + off_in : (w < n_in + n_mid ? w + off_mid : w + off_out)); };
  const uint32_t stride = gridDim.x * kBlock;
  for (uint32_t k0 = blockIdx.x * kBlock + threadIdx.x; k0 < G; k0 += stride * kGatherIlp) {
    uint32_t g[kGatherIlp];
    uint4 gt[kGatherIlp];
#pragma unroll
    for (int i = 0; i < kGatherIlp; ++i) {
      // a position past the end must NOT read order[]: its owner may already have shifted that entry to a global index
      uint32_t k = k0 + i * stride;
      g[i] = k < G ? (identity ? k : order[k]) : 0u;
    }
#pragma unroll
    for (int i = 0; i < kGatherIlp; ++i) gt[i] = identity ? ldg_stream(gates + g[i]) : __ldg(gates + g[i]);
#pragma unroll
    for (int i = 0; i < kGatherIlp; ++i) { gt[i].y = __ldg(wire + gt[i].y); gt[i].z = __ldg(wire + gt[i].z); gt[i].w = __ldg(wire + gt[i].w); }
#pragma unroll
    for (int i = 0; i < kGatherIlp; ++i) {
      uint32_t k = k0 + i * stride;
      if (k < G) {
        stg_stream(new_gates + k, make_uint4(gt[i].x, fix(gt[i].y), fix(gt[i].z), fix(gt[i].w)));
        if (order && gate_base) order[k] = g[i] + gate_base;  // each position is read and written by this thread only
      }
    }
  }
}

// Multi-GPU: local -> global wire ids / gate indices for one independently sorted component subtree.
__global__ void __launch_bounds__(kBlock) k_rebase(uint4* __restrict__ new_gates, uint32_t* __restrict__ order, uint32_t G, uint32_t n_in,
                                                   uint32_t n_mid, uint32_t off_in, uint32_t off_mid, uint32_t off_out, uint32_t gate_base) {
  auto fix = [&](uint32_t w) { return w < n_in ? w + off_in : (w < n_in + n_mid ? w + off_mid : w + off_out); };
  for (uint32_t k = blockIdx.x * kBlock + threadIdx.x; k < G; k += gridDim.x * kBlock) {
    uint4 g = new_gates[k];
    new_gates[k] = make_uint4(g.x, fix(g.y), fix(g.z), fix(g.w));
    if (order) order[k] += gate_base;
  }
}

// same, offsets derived from the all-gathered counts (rank-major {n_in, n_mid, n_out, G}) by every thread: <= 8 ranks x 4 words
__global__ void __launch_bounds__(kBlock) k_rebase_gathered(uint4* __restrict__ new_gates, uint32_t* __restrict__ order, uint32_t G,
                                                            const unsigned long long* __restrict__ counts, uint32_t rank, uint32_t world) {
  unsigned long long tot_in = 0, tot_mid = 0, pre_in = 0, pre_mid = 0, pre_out = 0, pre_g = 0;
  for (uint32_t r = 0; r < world; ++r) {
    unsigned long long a = counts[4 * r], b = counts[4 * r + 1], c = counts[4 * r + 2], g = counts[4 * r + 3];
    tot_in += a;
    tot_mid += b;
    if (r < rank) { pre_in += a; pre_mid += b; pre_out += c; pre_g += g; }
  }
  const uint32_t n_in = (uint32_t)counts[4 * rank], n_mid = (uint32_t)counts[4 * rank + 1];
  const uint32_t off_in = (uint32_t)pre_in, off_mid = (uint32_t)(tot_in + pre_mid) - n_in, off_out = (uint32_t)(tot_in + tot_mid + pre_out) - n_in - n_mid;
  const uint32_t gate_base = (uint32_t)pre_g;
  auto fix = [&](uint32_t w) { return w < n_in ? w + off_in : (w < n_in + n_mid ? w + off_mid : w + off_out); };
  for (uint32_t k = blockIdx.x * kBlock + threadIdx.x; k < G; k += gridDim.x * kBlock) {
    uint4 g = new_gates[k];
    new_gates[k] = make_uint4(g.x, fix(g.y), fix(g.z), fix(g.w));
    if (order) order[k] += gate_base;
  }
}

__global__ void __launch_bounds__(kBlock) k_rebase_ids_gathered(uint32_t* __restrict__ ids, uint64_t n, const unsigned long long* __restrict__ counts,
                                                                uint32_t rank, uint32_t world) {
  unsigned long long tot_in = 0, tot_mid = 0, pre_in = 0, pre_mid = 0, pre_out = 0;
  for (uint32_t r = 0; r < world; ++r) {
    tot_in += counts[4 * r];
    tot_mid += counts[4 * r + 1];
    if (r < rank) { pre_in += counts[4 * r]; pre_mid += counts[4 * r + 1]; pre_out += counts[4 * r + 2]; }
  }
  const uint32_t n_in = (uint32_t)counts[4 * rank], n_mid = (uint32_t)counts[4 * rank + 1];
  const uint32_t off_in = (uint32_t)pre_in, off_mid = (uint32_t)(tot_in + pre_mid) - n_in, off_out = (uint32_t)(tot_in + tot_mid + pre_out) - n_in - n_mid;
  for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kBlock) {
    uint32_t w = ids[i];
    if (w != kNone) ids[i] = w < n_in ? w + off_in : (w < n_in + n_mid ? w + off_mid : w + off_out);
  }
}

__global__ void __launch_bounds__(kBlock) k_rebase_map(uint32_t* __restrict__ wire, uint32_t n, uint32_t n_in, uint32_t n_mid, uint32_t off_in,
                                                       uint32_t off_mid, uint32_t off_out) {
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    uint32_t w = wire[i];
    if (w != kNone) wire[i] = w < n_in ? w + off_in : (w < n_in + n_mid ? w + off_mid : w + off_out);
  }
}

// ---------------------------------------------------------------------------------------------------
// host-side plumbing
// ---------------------------------------------------------------------------------------------------
int fail(c2a_handle* h, int status, const char* fmt, ...) {
  if (h) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    h->err = buf;
  }
  return status;
}

bool cuda_ok(c2a_handle* h, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  fail(h, C2A_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return false;
}

void slab_reset(c2a_handle* h) {
  h->slab_used = 0;
  h->slab_keep = 0;
  h->emitted.valid = false;
  h->emitted.prod1_valid = false;
  h->emitted.wire = nullptr;
}
void slab_reset_keep(c2a_handle* h) { h->slab_used = h->slab_keep; }

bool slab_reserve(c2a_handle* h, size_t bytes) {
  if (bytes <= h->slab_bytes) return true;
  char* old = h->slab;
  size_t want = bytes + bytes / 8 + (1u << 20);
  char* fresh = nullptr;
  cudaError_t e = cudaMalloc(&fresh, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (old && h->slab_keep == 0) {  // nothing to preserve: release first, then retry at the exact size
      cudaStreamSynchronize(h->stream);
      cudaFree(old);
      old = h->slab = nullptr;
      h->slab_bytes = 0;
    }
    e = cudaMalloc(&fresh, bytes);
    want = bytes;
  }
  if (e != cudaSuccess) {
    fail(h, C2A_ERR_NO_MEMORY, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    cudaGetLastError();
    return false;
  }
  if (old) {
    if (h->slab_keep) cudaMemcpyAsync(fresh, old, h->slab_keep, cudaMemcpyDeviceToDevice, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaFree(old);
  }
  h->slab = fresh;
  h->slab_bytes = want;
  return true;
}

void* slab_alloc(c2a_handle* h, size_t bytes) {
  size_t b = align256(bytes ? bytes : 1);
  if (h->slab_used + b > h->slab_bytes) return nullptr;
  void* p = h->slab + h->slab_used;
  h->slab_used += b;
  return p;
}

static cudaEvent_t next_event(c2a_handle* h) {
  if (h->ev_next == h->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->ev_pool.push_back(e);
  }
  return h->ev_pool[h->ev_next++];
}

void phases_clear(c2a_handle* h) {
  h->phases.clear();
  h->ev_next = 0;
  h->relax_fallback_rounds = 0;
}
// Every phase is also an NVTX range (the reference has `log` tracing only; here a timeline tool shows the pipeline
// phase by phase).  CUDA-event timing is optional on top of it.
void phase_begin(c2a_handle* h, const char* name) {
  nvtxRangePushA(name);
  h->nvtx_open++;
  if (!h->timing) return;
  if (!h->timing_only.empty() && h->timing_only != name) return;
  c2a_handle::Phase p{h->phase_prefix.empty() ? std::string(name) : h->phase_prefix + name, next_event(h), next_event(h), true};
  cudaEventRecord(p.a, h->stream);
  h->phases.push_back(p);
}
void phase_end(c2a_handle* h) {
  if (h->nvtx_open) { nvtxRangePop(); h->nvtx_open--; }
  if (!h->timing || h->phases.empty() || !h->phases.back().open) return;
  cudaEventRecord(h->phases.back().b, h->stream);
  h->phases.back().open = false;
}
void phases_collect(c2a_handle* h) {
  h->last_ms.clear();
  h->last_ms.push_back({"n_relax_fallback_rounds", (double)h->relax_fallback_rounds});  // a count, not a time (diagnostics)
  if (!h->timing || h->phases.empty()) return;
  std::map<std::string, double> acc;
  std::vector<std::string> names;
  for (auto& p : h->phases) {
    if (p.open) continue;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.a, p.b) != cudaSuccess) { cudaGetLastError(); continue; }
    if (!acc.count(p.name)) names.push_back(p.name);
    acc[p.name] += ms;
  }
  for (auto& n : names) h->last_ms.push_back({n, acc[n]});
  if (getenv("C2A_PHASE_TIMELINE")) {  // developer aid: where the stream idles between phases (start offset, duration, gap before)
    float prev_end = 0;
    for (auto& p : h->phases) {
      if (p.open) continue;
      float t0 = 0, d = 0;
      if (cudaEventElapsedTime(&t0, h->phases.front().a, p.a) != cudaSuccess || cudaEventElapsedTime(&d, p.a, p.b) != cudaSuccess) { cudaGetLastError(); continue; }
      fprintf(stderr, "[c2a timeline] %-28s start %8.3f ms  dur %7.3f ms  gap before %7.3f ms\n", p.name.c_str(), t0, d, t0 - prev_end);
      prev_end = t0 + d;
    }
  }
  float tot = 0;
  if (cudaEventElapsedTime(&tot, h->phases.front().a, h->phases.back().b) == cudaSuccess) h->last_ms.push_back({"total", tot});
  else cudaGetLastError();
}

int grid_for(c2a_handle* h, const void* kernel, int block, uint64_t n) {
  static std::unordered_map<const void*, int> per_sm;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  auto it = per_sm.find(kernel);
  int occ;
  if (it == per_sm.end()) {
    occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, 0) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 4; }
    per_sm[kernel] = occ;
  } else occ = it->second;
  uint64_t need = (n + block - 1) / block;
  uint64_t cap = (uint64_t)h->num_sms * occ;  // one full wave of resident CTAs, grid-stride inside
  uint64_t g = std::min<uint64_t>(std::max<uint64_t>(need, 1), cap);
  return (int)g;
}

#define LAUNCH(h, kernel, grid, block, ...)                           \
  do {                                                                \
    kernel<<<(grid), (block), 0, (h)->stream>>>(__VA_ARGS__);         \
    (h)->launches++;                                                  \
  } while (0)
#define LAUNCH_SIDE(h, kernel, grid, block, ...)                      \
  do {                                                                \
    kernel<<<(grid), (block), 0, (h)->stream2>>>(__VA_ARGS__);        \
    (h)->launches++;                                                  \
  } while (0)

static inline uint32_t scan_tiles(uint64_t n, int items) { return (uint32_t)((n + (uint64_t)kBlock * items - 1) / ((uint64_t)kBlock * items)); }

static inline uint32_t wire_bitmap_words(uint64_t n) { return (uint32_t)((3 * n + 31) / 32); }  // one bit per (sorted position, slot)

size_t sort_scratch_bytes(uint64_t n) {
  size_t b = 0;
  b += align256(4 * n);            // r
  b += align256(4 * (n + 1));      // size_off
  b += align256(n);                // state
  b += align256(4 * ((n + 31) / 32 + 1));  // inq
  b += 3 * align256(4 * n);        // q0 q1 heavy
  b += 6 * align256(4 * n);        // k_tree_blocks: par, anc x2, value x2, sz
  b += align256(8 * (size_t)(scan_tiles(n, kScanItems) + scan_tiles(wire_bitmap_words(n), kScanItems) + 2));  // tile_state: block-offset scan + bitmap scan
  b += 2 * align256(4 * ((size_t)wire_bitmap_words(n) + 4));  // first-appearance bitmap + its rank prefix
  b += align256(4 * S_COUNT);
  return b;
}

bool sort_scratch_carve(c2a_handle* h, uint64_t n, SortScratch* s) {
  s->r = (uint32_t*)slab_alloc(h, 4 * n);
  s->size_off = (uint32_t*)slab_alloc(h, 4 * (n + 1));
  s->state = (uint8_t*)slab_alloc(h, n);
  s->inq = (uint32_t*)slab_alloc(h, 4 * ((n + 31) / 32 + 1));
  s->q0 = (uint32_t*)slab_alloc(h, 4 * n);
  s->q1 = (uint32_t*)slab_alloc(h, 4 * n);
  s->heavy = (uint32_t*)slab_alloc(h, 4 * n);
  for (int i = 0; i < 6; ++i) s->tree[i] = (uint32_t*)slab_alloc(h, 4 * n);
  s->tile_state_bytes = 8 * (size_t)(scan_tiles(n, kScanItems) + scan_tiles(wire_bitmap_words(n), kScanItems) + 2);
  s->bitmap_words = wire_bitmap_words(n);
  s->bitmap = (uint32_t*)slab_alloc(h, 4 * ((size_t)s->bitmap_words + 4));
  s->bitmap_pre = (uint32_t*)slab_alloc(h, 4 * ((size_t)s->bitmap_words + 4));
  s->tile_state = (unsigned long long*)slab_alloc(h, s->tile_state_bytes);
  s->tile_state2 = s->tile_state ? s->tile_state + scan_tiles(n, kScanItems) + 1 : nullptr;
  s->scalars = (uint32_t*)slab_alloc(h, 4 * S_COUNT);
  return s->scalars != nullptr;
}

// scalars = 0, err = ~0; look-back tile states = 0.  Stream-ordered, no host wait.
void sort_scalars_reset(c2a_handle* h, const SortScratch& s) {
  cudaMemsetAsync(s.scalars, 0, 4 * S_COUNT, h->stream);
  cudaMemsetAsync(s.scalars + S_ERR_LO, 0xFF, 8, h->stream);
  cudaMemsetAsync(s.tile_state, 0, s.tile_state_bytes, h->stream);
  cudaMemsetAsync(s.bitmap, 0, 4 * ((size_t)s.bitmap_words + 4), h->stream);
}

// ---------------------------------------------------------------------------------------------------
// Exact reference order from dependency pairs (K5a-c), enqueued WITHOUT host round trips:
//   sort_enqueue_relax   scratch init, k_relax_loop (seeds + every relaxation round, cooperative)
//   sort_enqueue_emit    block sizes, offsets scan, size-1 blocks, per-tree DFS
// every kernel reads scalars[] to decide whether it has work (identity order / bad input / r[] not yet converged).
// (The relaxation runs to convergence inside one cooperative kernel, k_relax_loop: no host-drained continuation any more.)
// Precondition: scalars zeroed except S_FLAGS, S_ERR = ~0 (sort_scalars_reset + the deps kernel).
// ---------------------------------------------------------------------------------------------------
void sort_enqueue_relax(c2a_handle* h, const uint2* d_dep, uint32_t n, const SortScratch& s) {
  if (n == 0) return;
  uint32_t* sc = s.scalars;
  phase_begin(h, "init");
  LAUNCH(h, k_sort_init, grid_for(h, (const void*)k_sort_init, kBlock, (uint64_t)n + 1), kBlock, n, s.r, s.size_off, s.state, s.inq, sc);
  phase_end(h);
  phase_begin(h, "k_relax");
  {  // one cooperative launch: seeds + every relaxation round (the seeds live in heavy[] until k_roots reuses it)
    static int occ = 0;
    if (!occ && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k_relax_loop, kBlock, 0) != cudaSuccess || occ < 1)) { cudaGetLastError(); occ = 1; }
    const uint2* a_dep = d_dep;
    uint32_t a_n = n;
    const uint32_t* a_seeds = s.heavy;
    uint32_t *a_r = s.r, *a_inq = s.inq, *a_q0 = s.q0, *a_q1 = s.q1, *a_sc = sc;
    void* args[] = {&a_dep, &a_n, &a_seeds, &a_r, &a_inq, &a_q0, &a_q1, &a_sc};
    cudaLaunchCooperativeKernel((const void*)k_relax_loop, dim3(h->num_sms * std::min(occ, 4)), dim3(kBlock), args, 0, h->stream);
    h->launches++;
  }
  phase_end(h);
}

void sort_enqueue_emit(c2a_handle* h, const uint2* d_dep, uint32_t n, const SortScratch& s, uint32_t* d_order) {
  if (n == 0) return;
  uint32_t* sc = s.scalars;
  phase_begin(h, "k_sizes");
  LAUNCH(h, k_sizes, grid_for(h, (const void*)k_sizes, kBlock, n), kBlock, s.r, n, s.size_off, sc);
  phase_end(h);
  phase_begin(h, "k_scan_u32");
  LAUNCH(h, k_scan_u32_t<false>, scan_tiles(n, kScanItems), kBlock, s.size_off, s.size_off, n, s.tile_state, (uint32_t*)nullptr, sc, 1);
  phase_end(h);
  phase_begin(h, "k_roots");
  LAUNCH(h, k_roots, grid_for(h, (const void*)k_roots, kBlock, n), kBlock, s.r, s.size_off, n, d_dep, d_order, s.heavy, sc);
  phase_end(h);
  // heavy count is only known on the device: launch a grid sized for the worst case the hardware can hold
  phase_begin(h, "k_tree_blocks");
  {  // big tree-shaped blocks by pointer jumping (cooperative; exits at once when there is none)
    static int occ = 0;
    if (!occ && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k_tree_blocks, kBlock, 0) != cudaSuccess || occ < 1)) { cudaGetLastError(); occ = 1; }
    const uint2* a_dep = d_dep;
    const uint32_t *a_r = s.r, *a_off = s.size_off;
    uint32_t a_n = n;
    uint32_t* a_order = d_order;
    TreeArrays a_t{s.q0, s.q1, s.tree[0], s.tree[1], s.tree[2], s.tree[3], s.tree[4], s.tree[5]};
    uint32_t* a_sc = sc;
    void* args[] = {&a_dep, &a_r, &a_off, &a_n, &a_order, &a_t, &a_sc};
    cudaLaunchCooperativeKernel((const void*)k_tree_blocks, dim3(h->num_sms * std::min(occ, 4)), dim3(kBlock), args, 0, h->stream);
    h->launches++;
  }
  phase_end(h);
  phase_begin(h, "k_tree_dfs");
  LAUNCH(h, k_tree_dfs, h->num_sms * 8, 128, s.heavy, d_dep, s.r, s.size_off, s.state, d_order, (const uint32_t*)s.tree[5], sc);
  phase_end(h);
}

// status words -> reference error (topological_sort.rs:34-38)
int sort_status(c2a_handle* h, const uint32_t* host_scalars, uint64_t* err_index) {
  unsigned long long err = ((unsigned long long)host_scalars[S_ERR_HI] << 32) | host_scalars[S_ERR_LO];
  if (err != ~0ull) {
    uint64_t at = (uint32_t)err;
    if (err_index) *err_index = at;
    return fail(h, C2A_ERR_CYCLIC_DEPENDENCY, "detected at i=%llu", (unsigned long long)at);
  }
  return C2A_OK;
}

// ---------------------------------------------------------------------------------------------------
// build_circuit on device-resident gates
// ---------------------------------------------------------------------------------------------------
struct BuildPlan {
  uint64_t G;
  uint32_t node_bound, n_in, n_out;
  bool want_wire;  // wire numbering + gather wanted (build_circuit) or order only (topo_sort)
};

// the caller's host arrays (one synchronisation per call): the renumbered gates are copied chunk by chunk on the copy stream while the
// gather of the next chunk runs; `rest` enqueues the other copies on the main stream
struct HostCopyOut {
  uint4* new_gates_host;
  std::function<void()> rest;
};

static size_t core_scratch_bytes(const BuildPlan& p, size_t n_pairs) {  // n_pairs = n_in + n_out
  size_t b = 0;
  b += align256(4 * (size_t)p.node_bound);  // prod1
  b += align256(8 * p.G);                   // dep
  b += sort_scratch_bytes(p.G);
  b += align256(4 * p.G);                   // order (internal)
  b += align256(4 * n_pairs + 4);           // I/O node lists
  return b;
}

// Core: everything after the gates are on the device.  d_wire may be null (internal), d_order may be null.
// I/O node lists: either host arrays (staged through pinned memory) or, when d_io_ready != null, a device array holding
// the n_in input nodes followed by the n_out output nodes.  io_flags_dev (optional): a device word whose non-zero value
// means "an I/O signal could not be mapped" (set by the caller's mapping kernel); it is read with the final status.
// The whole pipeline is enqueued without intermediate host reads; ONE status read at the end.
static int build_core(c2a_handle* h, const BuildPlan& p, const uint4* d_gates, const uint32_t* in_nodes_host,
                      const uint32_t* out_nodes_host, uint32_t* d_order_user, uint32_t* d_wire, uint4* d_new_gates,
                      uint32_t* wire_count, uint64_t* err_index, bool* identity_out, const uint32_t* d_io_ready = nullptr,
                      const uint32_t* io_flags_dev = nullptr, const uint32_t* prod1_ready = nullptr,
                      const HostCopyOut* out = nullptr /* host arrays: copied out behind the pipeline, in front of the status read */) {
  cudaStream_t st = h->stream;
  const uint32_t G = (uint32_t)p.G;
  size_t n_pairs = p.want_wire ? (size_t)p.n_in + p.n_out : 0;
  if (!d_io_ready && (4 * n_pairs + 2048) > h->h_pinned_bytes) {
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    h->h_pinned_bytes = 4 * n_pairs + 8192;
    if (!cuda_ok(h, cudaHostAlloc((void**)&h->h_pinned, h->h_pinned_bytes, cudaHostAllocDefault), "cudaHostAlloc")) return C2A_ERR_CUDA;
  }

  // prod1_ready: the producer map of these gates already exists (the device emitter fills it while it resolves the gates)
  uint32_t* prod1 = prod1_ready ? const_cast<uint32_t*>(prod1_ready) : (uint32_t*)slab_alloc(h, 4 * (size_t)p.node_bound);
  uint2* dep = (uint2*)slab_alloc(h, 8 * p.G);
  SortScratch s;
  bool ok = sort_scratch_carve(h, p.G, &s);
  uint32_t* order_int = (uint32_t*)slab_alloc(h, 4 * p.G);
  uint32_t* io_nodes = d_io_ready ? const_cast<uint32_t*>(d_io_ready) : (uint32_t*)slab_alloc(h, 4 * n_pairs + 4);
  if (!ok || !prod1 || !dep || !order_int || !io_nodes) return fail(h, C2A_ERR_NO_MEMORY, "scratch slab exhausted");
  uint32_t* sc = s.scalars;
  uint32_t* hp = h->h_pinned;

  sort_scalars_reset(h, s);
  if (n_pairs && !d_io_ready) {
    uint32_t* stage = hp + 256;
    if (p.n_in) memcpy(stage, in_nodes_host, 4 * (size_t)p.n_in);
    if (p.n_out) memcpy(stage + p.n_in, out_nodes_host, 4 * (size_t)p.n_out);
    cudaMemcpyAsync(io_nodes, stage, 4 * n_pairs, cudaMemcpyHostToDevice, st);
  }

  // wire[] is first touched after the sort: its fill runs on the side stream next to K1 / K2 / K5 (several of which are
  // latency-bound) and the main stream joins in front of the first wire kernel
  bool wire_join = false;
  if (p.want_wire) {
    cudaEventRecord(h->ev_main, st);
    cudaStreamWaitEvent(h->stream2, h->ev_main, 0);
    if (p.node_bound) cudaMemsetAsync(d_wire, 0xFF, 4 * (size_t)p.node_bound, h->stream2);
    // ... and so do the I/O tags (inputs: their wire ids right away; outputs: "pending") - they depend on the fill only, not on the sort
    const uint32_t* t_in = io_nodes;
    const uint32_t* t_out = io_nodes + p.n_in;
    if (p.n_in) {
      LAUNCH_SIDE(h, k_io_set, grid_for(h, (const void*)k_io_set, kBlock, p.n_in), kBlock, t_in, p.n_in, p.node_bound, 0u, d_wire, sc);
      LAUNCH_SIDE(h, k_io_max, grid_for(h, (const void*)k_io_max, kBlock, p.n_in), kBlock, t_in, p.n_in, p.node_bound, 0u, (const uint32_t*)nullptr, d_wire);
    }
    if (p.n_out) LAUNCH_SIDE(h, k_io_set, grid_for(h, (const void*)k_io_set, kBlock, p.n_out), kBlock, t_out, p.n_out, p.node_bound, kOutPending, d_wire, sc);
    cudaEventRecord(h->ev_side, h->stream2);
    wire_join = true;
  }
  if (!prod1_ready) {
    phase_begin(h, "init");
    cudaMemsetAsync(prod1, 0, 4 * (size_t)p.node_bound, st);
    phase_end(h);
    phase_begin(h, "k_producer");
    if (G) LAUNCH(h, k_producer, grid_for(h, (const void*)k_producer, kBlock, G), kBlock, d_gates, G, p.node_bound, prod1, sc);
    phase_end(h);
  }
  phase_begin(h, "k_deps");
  if (G) LAUNCH(h, k_deps_t<false>, grid_for(h, (const void*)k_deps_t<false>, kBlock, G), kBlock, d_gates, G, p.node_bound, prod1, dep, s.heavy, (uint32_t*)nullptr, sc);
  phase_end(h);

  uint32_t* d_order = d_order_user ? d_order_user : order_int;
  sort_enqueue_relax(h, dep, G, s);

  const uint32_t ni = p.n_in, no = p.n_out;
  const uint32_t* d_out = io_nodes + ni;
  bool gates_copied = false;  // the copy stream took the renumbered gates, chunk by chunk
  // everything downstream of the relaxation; issued again after sort_drain() when the speculative rounds did not converge
  auto enqueue_tail = [&]() {
    sort_enqueue_emit(h, dep, G, s, d_order);
    if (d_order_user && G) {
      phase_begin(h, "k_iota");
      LAUNCH(h, k_iota_if_identity, grid_for(h, (const void*)k_iota_if_identity, kBlock, G), kBlock, d_order_user, G, sc);
      phase_end(h);
    }
    if (!p.want_wire) return;
    if (wire_join) { cudaStreamWaitEvent(st, h->ev_side, 0); wire_join = false; }  // wire[] filled, I/O nodes tagged
    if (G) {
      phase_begin(h, "k_wire_first");
      LAUNCH(h, k_wire_first, grid_for(h, (const void*)k_wire_first, kBlock, ((uint64_t)G + kWireIlp - 1) / kWireIlp), kBlock, d_gates, d_order, G, d_wire, sc);
      phase_end(h);
      phase_begin(h, "k_wire_mark");
      LAUNCH(h, k_wire_mark, grid_for(h, (const void*)k_wire_mark, kBlock, p.node_bound / 4 + 4), kBlock, d_wire, p.node_bound, s.bitmap, sc);
      phase_end(h);
      phase_begin(h, "k_scan_u32");
      LAUNCH(h, k_scan_u32_t<true>, scan_tiles(s.bitmap_words, kScanItems), kBlock, s.bitmap, s.bitmap_pre, s.bitmap_words, s.tile_state2, sc + S_NMID, sc, 2);
      phase_end(h);
      phase_begin(h, "k_wire_assign");
      LAUNCH(h, k_wire_assign, grid_for(h, (const void*)k_wire_assign, kBlock, p.node_bound / 4 + 4), kBlock, d_wire, p.node_bound, p.n_in, s.bitmap, s.bitmap_pre, sc);
      phase_end(h);
    }
    if (no) {
      LAUNCH(h, k_io_set, grid_for(h, (const void*)k_io_set, kBlock, no), kBlock, d_out, no, p.node_bound, 0u, d_wire, sc);
      LAUNCH(h, k_io_max, grid_for(h, (const void*)k_io_max, kBlock, no), kBlock, d_out, no, p.node_bound, p.n_in, (const uint32_t*)(sc + S_NMID), d_wire);
    }
    if (d_new_gates && G) {
      phase_begin(h, "k_gather");
      const bool chunked = out && out->new_gates_host && h->stream3 && G >= (1u << 22);
      const uint32_t nch = chunked ? 4u : 1u;
      for (uint32_t c = 0; c < nch; ++c) {
        const uint32_t lo = (uint32_t)((uint64_t)G * c / nch), hi = (uint32_t)((uint64_t)G * (c + 1) / nch);
        LAUNCH(h, k_gather, grid_for(h, (const void*)k_gather, kBlock, ((uint64_t)(hi - lo) + kGatherIlp - 1) / kGatherIlp), kBlock, d_gates, d_order, lo, hi - lo, d_wire,
               d_new_gates, sc);
        if (chunked) {
          cudaEventRecord(h->ev_main, st);
          cudaStreamWaitEvent(h->stream3, h->ev_main, 0);
          cudaMemcpyAsync(out->new_gates_host + lo, d_new_gates + lo, 16 * (size_t)(hi - lo), cudaMemcpyDeviceToHost, h->stream3);
        }
      }
      if (chunked) { cudaEventRecord(h->ev_copy, h->stream3); gates_copied = true; }
      phase_end(h);
    }
  };
  auto read_status = [&]() -> bool {
    if (!cuda_ok(h, cudaMemcpyAsync(hp + 64, sc, 4 * S_COUNT, cudaMemcpyDeviceToHost, st), "status copy")) return false;
    if (io_flags_dev) cudaMemcpyAsync(hp + 64 + S_COUNT, io_flags_dev, 4, cudaMemcpyDeviceToHost, st);
    if (!cuda_ok(h, cudaStreamSynchronize(st), "final sync")) return false;
    return cuda_ok(h, cudaGetLastError(), "kernel");
  };

  enqueue_tail();
  if (out) {
    phase_begin(h, "d2h");
    if (!gates_copied && out->new_gates_host && d_new_gates && G) cudaMemcpyAsync(out->new_gates_host, d_new_gates, 16 * (size_t)G, cudaMemcpyDeviceToHost, st);
    if (out->rest) out->rest();
    if (gates_copied) cudaStreamWaitEvent(st, h->ev_copy, 0);  // the status read below then covers the copy stream too
    phase_end(h);
  }
  if (!read_status()) return C2A_ERR_CUDA;
  const uint32_t* hs = hp + 64;
  if (io_flags_dev && hs[S_COUNT]) return fail(h, C2A_ERR_INVALID_ARGUMENT, "an input/output signal was never declared");
  if (hs[S_FLAGS] & F_BAD) return fail(h, C2A_ERR_INVALID_ARGUMENT, "a gate references a node id >= node_bound (%u)", p.node_bound);
  h->relax_fallback_rounds = hs[S_RELAX_ROUNDS] | (hs[S_ROBUST] ? 0x10000u : 0u);  // diagnostics: rounds run; bit 16 = pointer-jumping stage taken
  int stt = sort_status(h, hs, err_index);
  if (stt != C2A_OK) return stt;
  if (identity_out) *identity_out = !(hs[S_FLAGS] & (F_OOO | F_SELF));
  if (!p.want_wire) return C2A_OK;
  if (hs[S_FLAGS] & F_BAD_IO) return fail(h, C2A_ERR_INVALID_ARGUMENT, "an input/output node id is >= node_bound (%u)", p.node_bound);
  if (wire_count) *wire_count = p.n_in + hs[S_NMID] + p.n_out;
  return C2A_OK;
}

static int check_sizes(c2a_handle* h, uint64_t G, uint32_t node_bound) {
  if (!h) return C2A_ERR_INVALID_ARGUMENT;
  if (G > (1ull << 29)) return fail(h, C2A_ERR_INVALID_ARGUMENT, "G=%llu exceeds the 2^29 gate limit of the u32 position encoding", (unsigned long long)G);
  if (node_bound >= kOutPending) return fail(h, C2A_ERR_INVALID_ARGUMENT, "node_bound too large");
  if (!cuda_ok(h, cudaSetDevice(h->device), "cudaSetDevice")) return C2A_ERR_CUDA;
  return C2A_OK;
}

}  // namespace c2a

// =====================================================================================================
// C ABI
// =====================================================================================================
using namespace c2a;

extern "C" {

int c2a_abi_version(void) { return C2A_ABI_VERSION; }

int c2a_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int c2a_create(int device, c2a_handle** out) {
  if (!out) return C2A_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int n = c2a_device_count();
  if (n <= 0 || device < 0 || device >= n) return C2A_ERR_CUDA;  // no CPU fallback: fail loudly
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return C2A_ERR_CUDA; }
  c2a_handle* h = new c2a_handle();
  h->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); delete h; return C2A_ERR_CUDA; }
  if (cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_side2, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_counts, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    if (h->ev_side) cudaEventDestroy(h->ev_side);
    if (h->ev_side2) cudaEventDestroy(h->ev_side2);
    if (h->ev_counts) cudaEventDestroy(h->ev_counts);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    cudaStreamDestroy(h->stream);
    delete h;
    return C2A_ERR_CUDA;
  }
  if (cudaStreamCreateWithFlags(&h->stream3, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();  // (optional: without it the payload words are copied on the main stream)
    if (h->stream3) { cudaStreamDestroy(h->stream3); h->stream3 = nullptr; }
  }
  h->h_pinned_bytes = 1 << 16;
  if (cudaHostAlloc((void**)&h->h_emit_status, 256, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); h->h_emit_status = nullptr; }
  if (cudaHostAlloc((void**)&h->h_pinned, h->h_pinned_bytes, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    cudaEventDestroy(h->ev_side);
    cudaEventDestroy(h->ev_side2);
    cudaEventDestroy(h->ev_counts);
    cudaEventDestroy(h->ev_main);
    cudaStreamDestroy(h->stream2);
    cudaStreamDestroy(h->stream);
    delete h;
    return C2A_ERR_CUDA;
  }
  *out = h;
  return C2A_OK;
}

void c2a_destroy(c2a_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  if (h->slab) cudaFree(h->slab);
  if (h->ev_buf) cudaFree(h->ev_buf);
  if (h->cx_buf) cudaFree(h->cx_buf);
  if (h->fused_ctl) cudaFree(h->fused_ctl);
  if (h->plan_buf) cudaFree(h->plan_buf);
  if (h->io_buf) cudaFree(h->io_buf);
  if (h->cx_pinned) cudaFreeHost(h->cx_pinned);
  emit_drop_host(h);
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  if (h->h_emit_status) cudaFreeHost(h->h_emit_status);
  cudaStreamSynchronize(h->stream2);
  if (h->stream3) { cudaStreamSynchronize(h->stream3); cudaStreamDestroy(h->stream3); }
  if (h->ev_copy) cudaEventDestroy(h->ev_copy);
  cudaEventDestroy(h->ev_side);
  cudaEventDestroy(h->ev_side2);
  cudaEventDestroy(h->ev_counts);
  cudaEventDestroy(h->ev_main);
  cudaStreamDestroy(h->stream2);
  cudaStreamDestroy(h->stream);
  delete h;
}

const char* c2a_last_error(const c2a_handle* h) { return h ? h->err.c_str() : "null handle"; }
uint64_t c2a_kernel_launches(const c2a_handle* h) { return h ? h->launches : 0; }
double c2a_last_kernel_ms(const c2a_handle* h, const char* name) {
  if (!h || !name) return -1.0;
  for (auto& p : h->last_ms)
    if (p.first == name) return p.second;
  return -1.0;
}
void* c2a_stream(c2a_handle* h) { return h ? (void*)h->stream : nullptr; }  // cudaStream_t, for callers that time on it
void c2a_set_timing(c2a_handle* h, int on) { if (h) h->timing = on != 0; }
void c2a_set_timing_only(c2a_handle* h, const char* phase) { if (h) h->timing_only = phase ? phase : ""; }
// comma-separated "name=ms" list of the last call's phases (diagnostics)
const char* c2a_last_phases(c2a_handle* h) {
  static thread_local std::string s;
  s.clear();
  if (h) for (auto& p : h->last_ms) { char b[96]; snprintf(b, sizeof b, "%s%s=%.6f", s.empty() ? "" : ",", p.first.c_str(), p.second); s += b; }
  return s.c_str();
}

int c2a_build_circuit_device(c2a_handle* h, const c2a_gate* d_gates, uint64_t G, uint32_t node_bound, const uint32_t* input_nodes,
                             uint32_t n_in, const uint32_t* output_nodes, uint32_t n_out, uint32_t* d_order_out, uint32_t* d_wire_of_node,
                             c2a_gate* d_new_gates, uint32_t* wire_count, uint64_t* err_index) {
  int st = check_sizes(h, G, node_bound);
  if (st) return st;
  phases_clear(h);
  BuildPlan p{G, node_bound, n_in, n_out, true};
  slab_reset(h);
  size_t need = core_scratch_bytes(p, (size_t)n_in + n_out) + (d_wire_of_node ? 0 : align256(4 * (size_t)node_bound));
  if (!slab_reserve(h, need)) return C2A_ERR_NO_MEMORY;
  uint32_t* d_wire = d_wire_of_node ? d_wire_of_node : (uint32_t*)slab_alloc(h, 4 * (size_t)node_bound);
  st = build_core(h, p, (const uint4*)d_gates, input_nodes, output_nodes, d_order_out, d_wire, (uint4*)d_new_gates, wire_count, err_index, nullptr);
  cudaStreamSynchronize(h->stream);
  phases_collect(h);
  return st;
}

int c2a_rebase_wires_device(c2a_handle* h, c2a_gate* d_new_gates, uint32_t* d_order, uint64_t G, uint32_t n_in, uint32_t n_mid,
                            uint32_t off_in, uint32_t off_mid, uint32_t off_out, uint32_t gate_base) {
  int st = check_sizes(h, G, 1);
  if (st) return st;
  if (!d_new_gates) return fail(h, C2A_ERR_INVALID_ARGUMENT, "null argument");
  if (G) LAUNCH(h, k_rebase, grid_for(h, (const void*)k_rebase, kBlock, G), kBlock, (uint4*)d_new_gates, d_order, (uint32_t)G, n_in, n_mid, off_in, off_mid, off_out, gate_base);
  if (!cuda_ok(h, cudaStreamSynchronize(h->stream), "rebase")) return C2A_ERR_CUDA;
  return C2A_OK;
}

int c2a_rebase_wires_gathered_device(c2a_handle* h, c2a_gate* d_new_gates, uint32_t* d_order, uint64_t G, const uint64_t* d_counts, uint32_t rank,
                                     uint32_t world) {
  int st = check_sizes(h, G, 1);
  if (st) return st;
  if (!d_new_gates || !d_counts || rank >= world) return fail(h, C2A_ERR_INVALID_ARGUMENT, "bad argument");
  if (G) LAUNCH(h, k_rebase_gathered, grid_for(h, (const void*)k_rebase_gathered, kBlock, G), kBlock, (uint4*)d_new_gates, d_order, (uint32_t)G,
                (const unsigned long long*)d_counts, rank, world);
  return cuda_ok(h, cudaGetLastError(), "rebase launch") ? C2A_OK : C2A_ERR_CUDA;
}

int c2a_rebase_wire_ids_gathered_device(c2a_handle* h, uint32_t* d_wire_ids, uint64_t n, const uint64_t* d_counts, uint32_t rank, uint32_t world) {
  if (!h) return C2A_ERR_INVALID_ARGUMENT;
  if ((n && !d_wire_ids) || !d_counts || rank >= world) return fail(h, C2A_ERR_INVALID_ARGUMENT, "bad argument");
  if (!cuda_ok(h, cudaSetDevice(h->device), "cudaSetDevice")) return C2A_ERR_CUDA;
  if (n) LAUNCH(h, k_rebase_ids_gathered, grid_for(h, (const void*)k_rebase_ids_gathered, kBlock, n), kBlock, d_wire_ids, n, (const unsigned long long*)d_counts, rank, world);
  return cuda_ok(h, cudaGetLastError(), "rebase launch") ? C2A_OK : C2A_ERR_CUDA;
}

int c2a_rebase_wire_map_device(c2a_handle* h, uint32_t* d_wire_of_node, uint64_t n, uint32_t n_in, uint32_t n_mid, uint32_t off_in, uint32_t off_mid,
                               uint32_t off_out) {
  if (!h) return C2A_ERR_INVALID_ARGUMENT;
  if (n >= kOutPending) return fail(h, C2A_ERR_INVALID_ARGUMENT, "node_bound too large");
  if (!cuda_ok(h, cudaSetDevice(h->device), "cudaSetDevice")) return C2A_ERR_CUDA;
  if (n && !d_wire_of_node) return fail(h, C2A_ERR_INVALID_ARGUMENT, "null argument");
  if (n) LAUNCH(h, k_rebase_map, grid_for(h, (const void*)k_rebase_map, kBlock, n), kBlock, d_wire_of_node, (uint32_t)n, n_in, n_mid, off_in, off_mid, off_out);
  if (!cuda_ok(h, cudaStreamSynchronize(h->stream), "rebase map")) return C2A_ERR_CUDA;
  return C2A_OK;
}

int c2a_build_circuit(c2a_handle* h, const c2a_gate* gates, uint64_t G, uint32_t node_bound, const uint32_t* input_nodes, uint32_t n_in,
                      const uint32_t* output_nodes, uint32_t n_out, uint32_t* order_out, uint32_t* wire_of_node, c2a_gate* new_gates,
                      uint32_t* wire_count, uint64_t* err_index) {
  int st = check_sizes(h, G, node_bound);
  if (st) return st;
  phases_clear(h);
  BuildPlan p{G, node_bound, n_in, n_out, true};
  slab_reset(h);
  size_t need = core_scratch_bytes(p, (size_t)n_in + n_out) + align256(16 * G) * 2 + align256(4 * G) + align256(4 * (size_t)node_bound);
  if (!slab_reserve(h, need)) return C2A_ERR_NO_MEMORY;
  uint4* d_gates = (uint4*)slab_alloc(h, 16 * G);
  uint4* d_new = new_gates ? (uint4*)slab_alloc(h, 16 * G) : nullptr;
  uint32_t* d_order = order_out ? (uint32_t*)slab_alloc(h, 4 * G) : nullptr;
  uint32_t* d_wire = (uint32_t*)slab_alloc(h, 4 * (size_t)node_bound);
  cudaStream_t s = h->stream;
  phase_begin(h, "h2d");
  if (G && !cuda_ok(h, cudaMemcpyAsync(d_gates, gates, 16 * G, cudaMemcpyHostToDevice, s), "gates H2D")) return C2A_ERR_CUDA;
  phase_end(h);
  st = build_core(h, p, d_gates, input_nodes, output_nodes, d_order, d_wire, d_new, wire_count, err_index, nullptr);
  if (st == C2A_OK) {
    phase_begin(h, "d2h");
    if (order_out && G) cudaMemcpyAsync(order_out, d_order, 4 * G, cudaMemcpyDeviceToHost, s);
    if (wire_of_node && node_bound) cudaMemcpyAsync(wire_of_node, d_wire, 4 * (size_t)node_bound, cudaMemcpyDeviceToHost, s);
    if (new_gates && G) cudaMemcpyAsync(new_gates, d_new, 16 * G, cudaMemcpyDeviceToHost, s);
    phase_end(h);
    if (!cuda_ok(h, cudaStreamSynchronize(s), "D2H")) st = C2A_ERR_CUDA;
    // nodes that never appear keep an in-flight tag only if something went wrong; kNone otherwise
  } else {
    cudaStreamSynchronize(s);
  }
  phases_collect(h);
  return st;
}

int c2a_topo_sort(c2a_handle* h, const c2a_gate* gates, uint64_t G, uint32_t node_bound, uint32_t* order_out, uint64_t* err_index) {
  int st = check_sizes(h, G, node_bound);
  if (st) return st;
  if (!order_out && G) return fail(h, C2A_ERR_INVALID_ARGUMENT, "order_out is null");
  phases_clear(h);
  BuildPlan p{G, node_bound, 0, 0, false};
  slab_reset(h);
  size_t need = core_scratch_bytes(p, 0) + align256(16 * G) + align256(4 * G);
  if (!slab_reserve(h, need)) return C2A_ERR_NO_MEMORY;
  uint4* d_gates = (uint4*)slab_alloc(h, 16 * G);
  uint32_t* d_order = (uint32_t*)slab_alloc(h, 4 * G);
  cudaStream_t s = h->stream;
  phase_begin(h, "h2d");
  if (G && !cuda_ok(h, cudaMemcpyAsync(d_gates, gates, 16 * G, cudaMemcpyHostToDevice, s), "gates H2D")) return C2A_ERR_CUDA;
  phase_end(h);
  st = build_core(h, p, d_gates, nullptr, nullptr, d_order, nullptr, nullptr, nullptr, err_index, nullptr);
  if (st == C2A_OK && G) {
    phase_begin(h, "d2h");
    cudaMemcpyAsync(order_out, d_order, 4 * G, cudaMemcpyDeviceToHost, s);
    phase_end(h);
  }
  if (!cuda_ok(h, cudaStreamSynchronize(s), "topo_sort sync") && st == C2A_OK) st = C2A_ERR_CUDA;
  phases_collect(h);
  return st;
}

int c2a_topo_sort_deps(c2a_handle* h, uint64_t n, const uint64_t* dep_off, const uint32_t* dep_idx, uint32_t* order_out, uint64_t* err_index) {
  int st = check_sizes(h, n, 1);
  if (st) return st;
  if (n == 0) return C2A_OK;
  if (!dep_off || !order_out) return fail(h, C2A_ERR_INVALID_ARGUMENT, "null argument");
  uint64_t nnz = dep_off[n];
  if (nnz > 2 * n) return fail(h, C2A_ERR_INVALID_ARGUMENT, "more than 2 dependencies per item on average: not the reference's shape");
  phases_clear(h);
  slab_reset(h);
  size_t need = align256(8 * (n + 1)) + align256(4 * nnz + 4) + align256(8 * n) + sort_scratch_bytes(n) + align256(4 * n);
  if (!slab_reserve(h, need)) return C2A_ERR_NO_MEMORY;
  unsigned long long* d_off = (unsigned long long*)slab_alloc(h, 8 * (n + 1));
  uint32_t* d_idx = (uint32_t*)slab_alloc(h, 4 * nnz + 4);
  uint2* dep = (uint2*)slab_alloc(h, 8 * n);
  SortScratch s;
  bool ok = sort_scratch_carve(h, n, &s);
  uint32_t* d_order = (uint32_t*)slab_alloc(h, 4 * n);
  if (!ok || !d_off || !d_idx || !dep || !d_order) return fail(h, C2A_ERR_NO_MEMORY, "scratch slab exhausted");
  cudaStream_t stq = h->stream;
  uint32_t* hp = h->h_pinned;
  const uint32_t nn = (uint32_t)n;
  sort_scalars_reset(h, s);
  cudaMemcpyAsync(d_off, dep_off, 8 * (n + 1), cudaMemcpyHostToDevice, stq);
  if (nnz) cudaMemcpyAsync(d_idx, dep_idx, 4 * nnz, cudaMemcpyHostToDevice, stq);
  phase_begin(h, "k_deps");
  LAUNCH(h, k_deps_from_csr, grid_for(h, (const void*)k_deps_from_csr, kBlock, n), kBlock, d_off, d_idx, nn, dep, s.heavy, s.scalars);
  phase_end(h);
  sort_enqueue_relax(h, dep, nn, s);
  auto tail = [&]() {
    sort_enqueue_emit(h, dep, nn, s, d_order);
    LAUNCH(h, k_iota_if_identity, grid_for(h, (const void*)k_iota_if_identity, kBlock, n), kBlock, d_order, nn, s.scalars);
    cudaMemcpyAsync(hp + 64, s.scalars, 4 * S_COUNT, cudaMemcpyDeviceToHost, stq);
    return cuda_ok(h, cudaStreamSynchronize(stq), "sort sync") && cuda_ok(h, cudaGetLastError(), "sort kernels");
  };
  if (!tail()) return C2A_ERR_CUDA;
  const uint32_t* hs = hp + 64;
  if (hs[S_FLAGS] & F_BAD) return fail(h, C2A_ERR_INVALID_ARGUMENT, "dependency rows must have <= 2 entries with indices < n");
  st = sort_status(h, hs, err_index);
  if (st == C2A_OK) {
    cudaMemcpyAsync(order_out, d_order, 4 * n, cudaMemcpyDeviceToHost, stq);
    if (!cuda_ok(h, cudaStreamSynchronize(stq), "order D2H")) st = C2A_ERR_CUDA;
  }
  phases_collect(h);
  return st;
}

}  // extern "C"

#include "c2a_kahn.cuh"
#include "c2a_emit.cuh"
#include "c2a_fused.cuh"
#include "c2a_shard.cuh"
#include "c2a_eval.cuh"
