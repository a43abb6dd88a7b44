// c2a_fused.cuh — emit + build of a SMALL circuit in ONE cooperative kernel (include/c2a.h: c2a_compile_packed*).
//
// Below ~1 M gates the multi-kernel pipeline (c2a_emit.cuh + c2a_device.cu: ~35 launches and as many memsets / event operations) is bound by the
// host's enqueue rate: 0.27 ms per circuit whether it has 1 413 gates (BASELINE config 2) or 150 K (config 4).  Here the same
// phases run inside one persistent kernel, one CTA per SM (or fewer for tiny streams), separated by a grid barrier
// (one RED + an acquire poll, ~1 us) instead of a kernel boundary; data-dependent loops (Boruvka rounds, relaxation rounds) iterate on
// device-resident counters, so deep forward-edge cones need no host round trip either.  Everything is carved by upper bounds known
// before anything is counted (dense packed stream: G <= n_words/3, C <= n_words/2, S <= n - n_words/3), so the host enqueues
// copy-in, one memset, the kernel and the copy-out and synchronises ONCE.
//
// Semantics are exactly those of c2a_emit_packed_* followed by c2a_emitted_build_circuit* (src/compiler.rs:139-278, 321-464;
// src/topological_sort.rs:3-50): same node ids, same gate vector, same DFS order, same wire ids.  Streams on which the reference
// errors (or that the device emitter declines) are detected with the same flags and handed to the multi-kernel path, which
// replays them exactly; the fused kernel only ever commits results for valid streams.
//
// Phases (B = grid barrier):
//   F0  init scratch + per-tile kind counts                                   B
//   F1  tile-count scan (every CTA, in shared memory), rank + scatter events  B     compiler.rs:139-209
//   F2  Boruvka pick (round 1) + I/O list positions per signal                B     compiler.rs:213-278
//   F3  hook (round 1)                                                        B
//   F4  { pick B hook B } while live edges remain
//   F5  rank prefix of the effective-connection bitmap + class ids            B     compiler.rs:257
//   F6  node_of_signal, merge screens, I/O nodes tagged in wire[]             B     compiler.rs:157, 239-245, 392-395, 446-449
//   F7  gates -> node ids, producer map                                       B     compiler.rs:401-406
//   F8  deps, forward-edge seeds                                              B     compiler.rs:408-421
//   F9  (only if some dependency points forward) relax seed B { round B }* sizes B offsets (look-back) B roots B tree DFS B
//   F10 first appearances B bitmap marks B rank prefix (look-back) B wire ids B gather                compiler.rs:427-464
#pragma once

namespace c2a {

constexpr int kFusedBlock = 1024;
constexpr uint32_t kFusedMaxTiles = 4096;                 // 4 M events: the tile-count scan lives in shared memory
constexpr size_t kFusedSmem = 3 * 4 * (size_t)kFusedMaxTiles;  // dynamic shared memory of the kernel
constexpr uint32_t kInBase = 0x7FFFFFFEu;                 // wire[] tag of input list position i:  kInBase - i   (> kOutBase)
constexpr uint32_t kOutBase = 0x3FFFFFFFu;                // wire[] tag of output list position j: kOutBase - j  (outputs override inputs)
constexpr uint32_t kOutFloor = 0x20000000u;
// scalars of the fused kernel
enum { FS_G = 0, FS_C, FS_S, FS_EFLAGS, FS_NEFF, FS_ROUNDS, FS_BFLAGS, FS_SEEDN, FS_HEAVYN, FS_NMID, FS_ERR_LO = 10 /* ~(root << 32 | i), max = smallest; 0 = no cycle */, FS_ERR_HI = 11, FS_DONE = 12,
       FS_RELAX_ROUNDS = 13, FS_NCUT = 14, FS_QN = 16 /* 4 rotating queue counters */, FS_MSF = 20 /* 4 rotating: cand / cur counters */, FS_COUNT = 32 };

struct FusedParams {
  const uint8_t* kinds;
  const uint32_t* words;
  uint32_t n, n_words, tiles;
  uint32_t implicit;        // C2A_PACKED_IMPLICIT_OPERANDS: bit 7 of a gate / connection byte = one operand is the signal declared last
  const uint32_t* io_sigs;  // n_in input signal ids, then n_out output signal ids (device)
  uint32_t n_in, n_out;
  uint32_t G_ub, C_ub, S_ub, NB_ub;
  // emit scratch
  uint32_t *tile_g, *tile_c, *tile_i;
  uint2* sig_meta;
  uint4* egates;
  uint2* conn;
  uint32_t* conn_sb;
  uint8_t* outmark;
  uint32_t *parent, *best, *nidf, *eff, *effp, *cur;
  uint4* cand;
  uint32_t *in_idx1, *out_idx1;
  // resident results of the emit
  uint4* gates;
  uint32_t* nos;
  uint32_t* prod1;
  // build scratch
  uint2* dep;
  uint32_t *r, *size_off;
  uint8_t* state;
  uint32_t *inq, *q0, *q1, *heavy, *bitmap, *bitmap_pre;
  unsigned long long* agg;  // look-back aggregates: 2 scans x gridDim
  // results of the build: the caller's device arrays when the exact sizes fit their capacities (the kernel decides once it knows
  // them: FS_G after F1, the node bound after F5), else internal scratch of bound size (the host then reports the capacity)
  uint32_t* order_user;  uint32_t* order_int;   // order_int never null
  uint32_t* wire_user;   uint32_t* wire_int;    // wire_int: NB_ub entries, never null
  uint4* new_user;       uint4* new_int;        // both null when the renumbered gates are not wanted
  unsigned long long gates_cap;
  uint32_t wire_cap;
  // control block (double-buffered in the handle: this launch zeroes the other half for the next one - no memset per call)
  uint32_t* sc;          // FS_COUNT scalars, zero on entry
  unsigned int* bar;     // grid barrier counter, zero on entry
  uint32_t cluster;      // != 0: the whole grid is ONE thread-block cluster (<= 8 CTAs): the grid barrier is the hardware cluster barrier
  uint32_t* ctl_next;    // the other half, ctl_words u32
  uint32_t ctl_words;
  uint32_t* host_sc;     // pinned host memory (mapped): the scalars are stored here by CTA 0 before it exits - no D2H copy
  unsigned long long* trace;  // diagnostics (C2A_FUSED_TRACE=1): globaltimer of CTA 0 at every barrier exit, [0] = count; else null
};

// ---- memory access helpers.  Arrays written by other CTAs in an earlier phase are read through L2 (ld.cg): the barrier's fence
// invalidates the L1, this makes the phases independent of it.
template <typename T>
__device__ __forceinline__ T ldg2(const T* p) { return __ldcg(p); }

struct FusedCtx {
  unsigned int epoch;  // thread 0 only
  unsigned int nbar;
};

// Grid barrier: bar.sync orders the CTA's threads before thread 0's release (cumulative at gpu scope), the acquire poll orders
// thread 0 - and through the second bar.sync the whole CTA - after every other CTA's release.  No separate fences: each
// MEMBAR.GPU costs about as much as the barrier itself.
__device__ __forceinline__ void grid_bar(const FusedParams& P, FusedCtx& cx) {
  if (P.cluster) {
    // a stream of a few thousand events runs in <= 8 CTAs launched as one cluster: barrier.cluster (release / acquire at cluster
    // scope, which covers every thread of this grid) replaces the atomic round trip through L2 - ~0.3 us instead of ~1.2 us, and a
    // circuit of this size is nothing but ~17 barriers
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (P.trace && blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (++cx.nbar < 120) { P.trace[cx.nbar] = t; P.trace[0] = cx.nbar; }
    }
    return;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    cx.epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(P.bar) : "memory");
    unsigned int v;
    uint32_t spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(P.bar) : "memory");
      if (++spins > (1u << 26)) __trap();  // a CTA never arrived: fail loudly instead of hanging
    } while (v < cx.epoch);
    if (P.trace && blockIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (++cx.nbar < 120) { P.trace[cx.nbar] = t; P.trace[0] = cx.nbar; }
    }
  }
  __syncthreads();
}

// diagnostics: an extra timeline entry inside a phase (all threads of the CTA must reach it)
__device__ __forceinline__ void trace_mark(const FusedParams& P, FusedCtx& cx) {
  if (!P.trace) return;
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (++cx.nbar < 120) { P.trace[cx.nbar] = t | (1ull << 63); P.trace[0] = cx.nbar; }
  }
}

#define FUSED_FOR(i, n) for (uint32_t i = blockIdx.x * kFusedBlock + threadIdx.x; i < (n); i += gridDim.x * kFusedBlock)

// block-wide exclusive scan of one value per thread (kFusedBlock threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp /* 33 u32 */, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = warp_incl_scan(v, lane);
  __syncthreads();  // s_warp may still be read from a previous call
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane];
    uint32_t wi = warp_incl_scan(w, lane);
    s_warp[lane] = wi - w;
    if (lane == 31) s_warp[32] = wi;
  }
  __syncthreads();
  if (total) *total = s_warp[32];
  return s_warp[warp] + incl - v;
}

// Grid-wide exclusive scan: CTA b owns the contiguous chunk b of the sequence, stores its aggregate, one grid barrier, then every
// CTA sums the aggregates before it (one load per thread) and scans its chunk.  f(i) = the i-th value.  Writes dst[i] = exclusive
// prefix; returns the grid total to every thread.  (A barrier-free variant - tagged aggregates polled by every CTA - was measured
// at 3-10 us per scan: 148 CTAs x 148 polling threads slow the very stores they wait for.)
template <typename F>
__device__ __forceinline__ uint32_t grid_scan(const FusedParams& P, FusedCtx& cx, uint32_t n, F f, uint32_t* dst, unsigned long long* agg, uint32_t* s_warp) {
  __shared__ uint32_t s_pre, s_tot;
  const uint32_t per = (n + gridDim.x - 1) / gridDim.x;
  const uint32_t lo = min(n, blockIdx.x * per), hi = min(n, lo + per);
  uint32_t sum = 0;
  for (uint32_t i = lo + threadIdx.x; i < hi; i += kFusedBlock) sum += f(i);
  uint32_t tot = 0;
  block_excl_scan(sum, s_warp, &tot);
  if (threadIdx.x == 0) agg[blockIdx.x] = tot;
  grid_bar(P, cx);
  {
    uint32_t pre = 0, all = 0;
    if (threadIdx.x < gridDim.x) {
      all = (uint32_t)__ldcg(agg + threadIdx.x);
      if (threadIdx.x < blockIdx.x) pre = all;
    }
    uint32_t tp = 0, ta = 0;
    block_excl_scan(pre, s_warp, &tp);
    block_excl_scan(all, s_warp, &ta);
    if (threadIdx.x == 0) { s_pre = tp; s_tot = ta; }
  }
  __syncthreads();
  uint32_t carry = s_pre;
  for (uint32_t base = lo; base < hi; base += kFusedBlock) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < hi ? f(i) : 0u, t = 0;
    uint32_t ex = block_excl_scan(v, s_warp, &t);
    if (i < hi) dst[i] = carry + ex;
    carry += t;
  }
  return s_tot;
}

__device__ __forceinline__ uint32_t fused_find(uint32_t* parent, uint32_t x) {
  uint32_t r = x, p;
  while ((p = ldg2(parent + r)) != r) r = p;
  while (x != r) { p = ldg2(parent + x); if (p != r) parent[x] = r; x = p; }
  return r;
}
__device__ __forceinline__ void fused_red_min(uint32_t* p, uint32_t v) {
  if (ldg2(p) > v) atomicMin(p, v);
}

// K5a inside the fused kernel (same protocol as relax_from / enqueue in c2a_device.cu; no __restrict__: dep[] and r[] were written
// earlier in this very kernel)
__device__ __forceinline__ void fused_enqueue(uint32_t x, uint32_t* inq, uint32_t* q, uint32_t* qn) {
  uint32_t bit = 1u << (x & 31);
  __threadfence();
  uint32_t old = atomicOr(inq + (x >> 5), bit);
  if (!(old & bit)) q[atomicAdd(qn, 1u)] = x;
}
// (walks are bounded like in k_relax_loop; a stream whose DAG turns out to be deep leaves the fused kernel - EF_DEEP - and goes
//  through the multi-kernel path, which has the pointer-jumping stages)
constexpr uint32_t EF_DEEP = 512;
// (the rh side of a lowered gate goes on a per-thread stack, as in relax_capped of c2a_device.cu: fewer rounds, and a round costs
//  three grid barriers here)
__device__ __forceinline__ void fused_relax_from(uint32_t cur, uint32_t val, const uint2* dep, uint32_t* r, uint32_t* inq, uint32_t* q, uint32_t* qn,
                                                 uint32_t cap, uint32_t* ncut) {
  uint32_t stk[kRelaxStack];
  int sp = 0;
  for (uint32_t hop = 0;; ++hop) {
    if (cur == kNone) {
      if (!sp) return;
      cur = stk[--sp];
    }
    if (hop == cap) {
      fused_enqueue(cur, inq, q, qn);
      while (sp) fused_enqueue(stk[--sp], inq, q, qn);
      atomicAdd(ncut, 1u);
      return;
    }
    uint2 d = ldg2(dep + cur);
    uint32_t nxt = kNone;
    if (d.y != kNone && d.y != d.x && val < __ldcg(r + d.y)) {
      if (val < atomicMin(r + d.y, val)) {
        if (sp < kRelaxStack) stk[sp++] = d.y;
        else fused_enqueue(d.y, inq, q, qn);
      }
    }
    if (d.x != kNone && val < __ldcg(r + d.x)) {
      if (val < atomicMin(r + d.x, val)) nxt = d.x;
    }
    cur = nxt;
  }
}

// CTA 0 hands the scalars to the host through mapped pinned memory (every exit of the kernel follows a grid barrier, or writes
// of CTA 0's own thread 0, so they are complete)
__device__ __forceinline__ void fused_publish(const FusedParams& P) {
  if (blockIdx.x != 0) return;
  __syncthreads();
  if (threadIdx.x < FS_COUNT) {
    volatile uint32_t* hs = P.host_sc;
    hs[threadIdx.x] = __ldcg(P.sc + threadIdx.x);
    __threadfence_system();
  }
}

__global__ void __launch_bounds__(kFusedBlock, 1) k_fused_compile(const FusedParams P) {
  extern __shared__ uint32_t s_dyn[];  // 3 x kFusedMaxTiles: exclusive tile prefixes of the gates, connections, implicit-operand events
  uint32_t* const s_tg = s_dyn;
  uint32_t* const s_tc = s_dyn + kFusedMaxTiles;
  uint32_t* const s_ti = s_dyn + 2 * kFusedMaxTiles;
  __shared__ uint32_t s_warp[33], s_wg[33], s_wc[33], s_wi[33];
  FusedCtx cx;
  cx.epoch = 0;
  cx.nbar = 0;
  if (P.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.trace[127] = t;  // kernel start
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t* sc = P.sc;
  const uint32_t n = P.n;

  // ================= F0: init + per-tile counts (one warp per 1024-event tile) =================
  FUSED_FOR(i, P.S_ub) { P.parent[i] = i; P.best[i] = 0xFFFFFFFFu; P.nidf[i] = 0; P.outmark[i] = 0; P.in_idx1[i] = 0; P.out_idx1[i] = 0; }
  FUSED_FOR(i, P.C_ub / 32 + 4) P.eff[i] = 0;
  FUSED_FOR(i, P.NB_ub) P.prod1[i] = 0;
  FUSED_FOR(i, P.ctl_words) P.ctl_next[i] = 0;
  FUSED_FOR(i, P.G_ub + 1) { P.size_off[i] = 0; if (i < P.G_ub) { P.r[i] = i; P.state[i] = 0; } }
  FUSED_FOR(i, (P.G_ub + 31) / 32 + 1) P.inq[i] = 0;
  FUSED_FOR(i, (3 * P.G_ub + 31) / 32 + 4) P.bitmap[i] = 0;
  trace_mark(P, cx);
  {
    uint32_t f = 0;
    const uint32_t nwarps = gridDim.x * (kFusedBlock / 32);
    const bool aligned = !(reinterpret_cast<uintptr_t>(P.kinds) & 15);
    for (uint32_t tile = blockIdx.x * (kFusedBlock / 32) + warp; tile < P.tiles; tile += nwarps) {
      const uint32_t tbase = tile * kEvTile;
      uint32_t g = 0, c = 0, im = 0;
      const uint32_t opmask = P.implicit ? 31u : 63u;
      uint32_t wd[8];
      if (aligned && tbase + kEvTile <= n) {  // two 128-bit loads per lane, both in flight
        const uint4 x = *(reinterpret_cast<const uint4*>(P.kinds + tbase) + lane), y = *(reinterpret_cast<const uint4*>(P.kinds + tbase + 512) + lane);
        wd[0] = x.x; wd[1] = x.y; wd[2] = x.z; wd[3] = x.w; wd[4] = y.x; wd[5] = y.y; wd[6] = y.z; wd[7] = y.w;
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          wd[q] = 0;  // padding = SIGNAL with op 0: neither counted nor flagged
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t i = tbase + (q < 4 ? 0 : 512) + lane * 16 + (q & 3) * 4 + j;
            if (i < n) wd[q] |= (uint32_t)P.kinds[i] << (8 * j);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t lo = wd[q] & 0x01010101u, hi = (wd[q] >> 1) & 0x01010101u;  // kind bit 0 / bit 1 of each byte
        g += __popc(hi & ~lo);
        c += __popc(hi & lo);
        if (P.implicit) {
          const uint32_t b7 = (wd[q] >> 7) & 0x01010101u;
          im += __popc(b7 & hi);
          if (b7 & ~hi) f |= EF_BAD_KIND;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t kb = (wd[q] >> (8 * j)) & 0xFFu, op = (kb >> 2) & opmask;
          if ((kb & 3u) == C2A_EV_GATE) { if (op >= C2A_GATE_TYPE_COUNT) f |= EF_BAD_OP; }
          else if (op) f |= EF_BAD_KIND;
        }
      }
      g = warp_sum(g);
      c = warp_sum(c);
      im = warp_sum(im);
      if (lane == 0) { P.tile_g[tile] = g; P.tile_c[tile] = c; P.tile_i[tile] = im; }
    }
    f = warp_or(f);
    if (lane == 0 && f) atomicOr(sc + FS_EFLAGS, f);
  }
  grid_bar(P, cx);

  // ================= F1: tile-count scan in shared memory (every CTA), then rank + scatter =================
  uint32_t G, C, S, I;
  {
    constexpr uint32_t per = kFusedMaxTiles / kFusedBlock;  // 4 tiles per thread
    uint32_t lg[per], lc[per], li[per], sg = 0, scn = 0, si = 0;
#pragma unroll
    for (uint32_t j = 0; j < per; ++j) {
      uint32_t t = threadIdx.x * per + j;
      lg[j] = t < P.tiles ? ldg2(P.tile_g + t) : 0u;
      lc[j] = t < P.tiles ? ldg2(P.tile_c + t) : 0u;
      li[j] = t < P.tiles ? ldg2(P.tile_i + t) : 0u;
      sg += lg[j];
      scn += lc[j];
      si += li[j];
    }
    uint32_t totg = 0, totc = 0, toti = 0;
    uint32_t eg = block_excl_scan(sg, s_warp, &totg);
    uint32_t ec = block_excl_scan(scn, s_warp, &totc);
    uint32_t ei = block_excl_scan(si, s_warp, &toti);
#pragma unroll
    for (uint32_t j = 0; j < per; ++j) {
      uint32_t t = threadIdx.x * per + j;
      s_tg[t] = eg; s_tc[t] = ec; s_ti[t] = ei;
      eg += lg[j]; ec += lc[j]; ei += li[j];
    }
    __syncthreads();
    G = totg; C = totc; S = n - G - C; I = toti;
  }
  trace_mark(P, cx);
  // results go straight into the caller's arrays when they hold G entries (every CTA derives the same verdict)
  const bool gates_fit = (unsigned long long)G <= P.gates_cap;
  uint32_t* const order = (P.order_user && gates_fit) ? P.order_user : P.order_int;
  uint4* const new_gates = (P.new_user && gates_fit) ? P.new_user : P.new_int;
  uint32_t* wire = P.wire_int;  // chosen in F5, when the node bound is known
  // the counts decide everything downstream: every CTA derives the same verdict from the same numbers
  const bool words_ok = (unsigned long long)P.n_words + I == 3ull * G + 2ull * C;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc[FS_G] = G; sc[FS_C] = C; sc[FS_S] = S;
    if (!words_ok) atomicOr(sc + FS_EFLAGS, (uint32_t)EF_CAP);  // n_words does not match the kinds: the host reports it
  }
  if (!words_ok || (ldg2(sc + FS_EFLAGS) & (EF_BAD_KIND | EF_BAD_OP))) { fused_publish(P); return; }  // uniform: flags were complete at the barrier
  {
    uint32_t f = 0;
    for (uint32_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
      const uint32_t tbase = tile * kEvTile, k = threadIdx.x, i = tbase + k;
      const uint32_t kb = i < n ? (uint32_t)P.kinds[i] : 0x100u;
      const bool is_g = kb < 0x100u && (kb & 3u) == C2A_EV_GATE, is_c = kb < 0x100u && (kb & 3u) == C2A_EV_CONNECT;
      const bool is_i = P.implicit && (is_g || is_c) && (kb & 0x80u);  // one operand is the signal declared last: no payload word for it
      const uint32_t gm = __ballot_sync(0xFFFFFFFFu, is_g), cm = __ballot_sync(0xFFFFFFFFu, is_c), im = __ballot_sync(0xFFFFFFFFu, is_i);
      __syncthreads();  // s_wg / s_wc of the previous tile
      if (lane == 0) { s_wg[warp] = __popc(gm); s_wc[warp] = __popc(cm); s_wi[warp] = __popc(im); }
      __syncthreads();
      if (warp == 0) {
        uint32_t a = s_wg[lane], b = s_wc[lane], c3 = s_wi[lane];
        uint32_t ai = warp_incl_scan(a, lane), bi = warp_incl_scan(b, lane), ci = warp_incl_scan(c3, lane);
        __syncwarp();
        s_wg[lane] = ai - a; s_wc[lane] = bi - b; s_wi[lane] = ci - c3;
      }
      __syncthreads();
      const uint32_t dg = s_wg[warp] + __popc(gm & lt), dc = s_wc[warp] + __popc(cm & lt), di = s_wi[warp] + __popc(im & lt), ds = k - dg - dc;
      const uint32_t g0 = s_tg[tile], c0 = s_tc[tile], s0 = tbase - g0 - c0;
      const uint32_t before = s0 + ds;  // signals declared before this event: dense ids => "declared before use" is id < before
      const unsigned long long w = 3ull * g0 + 2ull * c0 - s_ti[tile] + 3u * dg + 2u * dc - di;
      if (is_g) {
        uint4 gt = make_uint4(P.implicit ? (kb >> 2) & 31u : kb >> 2, P.words[w], P.words[w + 1], is_i ? before - 1u : P.words[w + 2]);
        if (gt.y < before && gt.z < before && gt.w < before) P.outmark[gt.w] = 1;  // compiler.rs:201
        else { f |= EF_UNKNOWN_REF; gt.y = gt.z = gt.w = 0; }
        P.egates[g0 + dg] = gt;
      } else if (is_c) {
        uint2 ab = is_i ? make_uint2(before - 1u, P.words[w]) : make_uint2(P.words[w], P.words[w + 1]);
        if (!(ab.x < before && ab.y < before)) { f |= EF_UNKNOWN_REF; ab = make_uint2(0, 0); }
        P.conn[c0 + dc] = ab;
        P.conn_sb[c0 + dc] = before;
      } else if (i < n) {
        P.sig_meta[before] = make_uint2(before | ((kb & 3u) == C2A_EV_SIGNAL_CONST ? 0x80000000u : 0u), c0 + dc);
      }
    }
    f = warp_or(f);
    if (lane == 0 && f) atomicOr(sc + FS_EFLAGS, f);
  }
  grid_bar(P, cx);

  // ================= F2: Boruvka round 1 pick (classes are the signals) + I/O list positions =================
  uint32_t rounds = 0;
  {
    const uint32_t tag = (6u - (rounds % 7u)) << 29;
    FUSED_FOR(e, C) {
      uint2 ab = ldg2(P.conn + e);
      if (ab.x != ab.y) { fused_red_min(P.best + ab.x, tag | e); fused_red_min(P.best + ab.y, tag | e); }
    }
    bool bad = false;
    FUSED_FOR(i, P.n_in + P.n_out) {
      uint32_t s = P.io_sigs[i];
      if (s >= S) bad = true;
      else if (i < P.n_in) atomicMax(P.in_idx1 + s, i + 1);
      else atomicMax(P.out_idx1 + s, i - P.n_in + 1);
    }
    if (bad) atomicOr(sc + FS_EFLAGS, (uint32_t)EF_BAD_IO);
  }
  grid_bar(P, cx);
  // ================= F3: hook round 1 =================
  {
    const uint32_t tag = (6u - (rounds % 7u)) << 29;
    const uint32_t Cw = (C + 31) & ~31u;
    FUSED_FOR(i, Cw) {
      bool keep = false, eff = false;
      if (i < C) {
        uint2 ab = ldg2(P.conn + i);
        if (ab.x != ab.y) {
          uint32_t val = tag | i;
          bool bu = ldg2(P.best + ab.x) == val, bv = ldg2(P.best + ab.y) == val;
          if (bu && bv) P.parent[max(ab.x, ab.y)] = min(ab.x, ab.y);
          else if (bu) P.parent[ab.x] = ab.y;
          else if (bv) P.parent[ab.y] = ab.x;
          else keep = true;
          eff = !keep;
        }
      }
      uint32_t m = __ballot_sync(0xFFFFFFFFu, eff);
      if (lane == 0 && m) P.eff[i >> 5] = m;  // the warp's 32 connections are exactly one bitmap word
      warp_append(keep, i, P.cur, sc + FS_MSF + 1);
    }
    if (C && blockIdx.x == 0 && threadIdx.x == 0) sc[FS_ROUNDS] = 1;
    ++rounds;
  }
  grid_bar(P, cx);
  // ================= F4: further rounds while undecided edges remain =================
  // counters rotate: round j reads cur count FS_MSF + (2j-1)%4, writes cand count FS_MSF + (2j)%4 and cur count FS_MSF + (2j+1)%4
  if (C) {
    uint32_t n_cur = ldg2(sc + FS_MSF + 1);
    uint32_t slot_cur = 1;
    while (n_cur) {
      const uint32_t slot_cand = (slot_cur + 1) & 3, slot_next = (slot_cur + 2) & 3;
      if ((rounds % 7u) == 0) {  // the tags wrapped: forget the old minima (one extra phase every 7 rounds)
        FUSED_FOR(i, S) P.best[i] = 0xFFFFFFFFu;
        grid_bar(P, cx);
      }
      const uint32_t tag = (6u - (rounds % 7u)) << 29;
      if (blockIdx.x == 0 && threadIdx.x == 0) sc[FS_MSF + slot_next] = 0;
      const uint32_t nw = (n_cur + 31) & ~31u;
      FUSED_FOR(i, nw) {
        bool keep = false;
        uint4 out = make_uint4(0, 0, 0, 0);
        if (i < n_cur) {
          uint32_t e = ldg2(P.cur + i);
          uint2 ab = ldg2(P.conn + e);
          uint32_t cu = fused_find(P.parent, ab.x), cv = fused_find(P.parent, ab.y);
          if (cu != cv) { fused_red_min(P.best + cu, tag | e); fused_red_min(P.best + cv, tag | e); keep = true; out = make_uint4(e, cu, cv, 0); }
        }
        warp_append(keep, out, P.cand, sc + FS_MSF + slot_cand);
      }
      grid_bar(P, cx);
      const uint32_t n_cand = ldg2(sc + FS_MSF + slot_cand);
      if (!n_cand) break;  // every remaining edge became internal
      if (blockIdx.x == 0 && threadIdx.x == 0) { sc[FS_MSF + ((slot_next + 1) & 3)] = 0; sc[FS_ROUNDS] = rounds + 1; }
      const uint32_t nc = (n_cand + 31) & ~31u;
      FUSED_FOR(i, nc) {
        bool keep = false;
        uint32_t e = 0;
        if (i < n_cand) {
          uint4 c = ldg2(P.cand + i);
          e = c.x;
          uint32_t val = tag | e;
          bool bu = ldg2(P.best + c.y) == val, bv = ldg2(P.best + c.z) == val;
          if (bu && bv) P.parent[max(c.y, c.z)] = min(c.y, c.z);
          else if (bu) P.parent[c.y] = c.z;
          else if (bv) P.parent[c.z] = c.y;
          else keep = true;
          if (!keep) atomicOr(P.eff + (e >> 5), 1u << (e & 31));
        }
        warp_append(keep, e, P.cur, sc + FS_MSF + slot_next);
      }
      ++rounds;
      grid_bar(P, cx);
      n_cur = ldg2(sc + FS_MSF + slot_next);
      slot_cur = slot_next;
      if (rounds > 96) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(sc + FS_EFLAGS, (uint32_t)EF_CAP); break; }
    }
  }
  // ================= F5: rank prefix of the effective-connection bitmap + class ids (last effective connection) =================
  uint32_t n_eff = 0;
  {
    const uint32_t W = C / 32 + 1;
    auto f = [&](uint32_t w) { return (uint32_t)__popc(ldg2(P.eff + w)); };
    n_eff = grid_scan(P, cx, W, f, P.effp, P.agg, s_warp);
    trace_mark(P, cx);
    __syncthreads();  // this CTA's chunk of effp is complete (it is the chunk whose connections it handles next)
    const uint32_t per = (W + gridDim.x - 1) / gridDim.x;
    const uint32_t wlo = min(W, blockIdx.x * per), whi = min(W, wlo + per);
    const uint32_t chi = min(C, whi * 32);
    for (uint32_t c = wlo * 32 + threadIdx.x; c < chi; c += kFusedBlock) {
      const uint32_t wd = ldg2(P.eff + (c >> 5));
      if (!((wd >> (c & 31)) & 1u)) continue;  // not effective: no id consumed (compiler.rs:235-237)
      const uint32_t id = ldg2(P.conn_sb + c) + P.effp[c >> 5] + __popc(wd & ((1u << (c & 31)) - 1u)) + 1u;  // compiler.rs:257
      const uint32_t root = fused_find(P.parent, ldg2(P.conn + c).x);
      if ((ldg2(P.nidf + root) & kNidMask) < id) atomicMax(P.nidf + root, id);  // (a class may hold most connections: skip what cannot raise the word)
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) sc[FS_NEFF] = n_eff;
    // the node bound is known: pick the wire map (the caller's when it holds S + n_eff + 1 entries) and clear it
    const uint32_t NBk = S + n_eff + 1;
    wire = (P.wire_user && NBk <= P.wire_cap) ? P.wire_user : P.wire_int;
    FUSED_FOR(i, NBk) wire[i] = kNone;
  }
  grid_bar(P, cx);
  // ================= F6: node_of_signal, merge screens, I/O nodes =================
  {
    uint32_t f = 0;
    FUSED_FOR(s, S) {
      const uint2 m = ldg2(P.sig_meta + s);
      const uint32_t r = fused_find(P.parent, s);
      uint32_t node = ldg2(P.nidf + r) & kNidMask;
      if (node == 0) {
        const uint32_t c = m.y;
        node = (m.x & 0x7FFFFFFFu) + 1u + ldg2(P.effp + (c >> 5)) + __popc(ldg2(P.eff + (c >> 5)) & ((1u << (c & 31)) - 1u));  // compiler.rs:157
      } else {  // merged class: at most one constant and one gate output may meet in it (compiler.rs:239-245)
        if (m.x & 0x80000000u) { if (atomicOr(P.nidf + r, kHasConst) & kHasConst) f |= EF_CONST_CONST; }
        if (ldg2(P.outmark + s)) { if (atomicOr(P.nidf + r, kHasOut) & kHasOut) f |= EF_OUT_OUT; }
      }
      P.nos[s] = node;
      // compiler.rs:392-395 / 446-449: list order, a node listed twice keeps the LAST position; an output tag beats an input tag
      const uint32_t i1 = ldg2(P.in_idx1 + s), o1 = ldg2(P.out_idx1 + s);
      if (i1) atomicMin(wire + node, kInBase - (i1 - 1));
      if (o1) atomicMin(wire + node, kOutBase - (o1 - 1));
    }
    f = warp_or(f);
    if (lane == 0 && f) atomicOr(sc + FS_EFLAGS, f);
  }
  grid_bar(P, cx);
  if (ldg2(sc + FS_EFLAGS)) { fused_publish(P); return; }  // the reference errors on this stream (or it needs the exact host replay): nothing is committed
  const uint32_t NB = S + n_eff + 1;  // node_bound
  // ================= F7: gates -> node ids + producer map (K1) =================
  FUSED_FOR(g, G) {
    const uint4 e = ldg2(P.egates + g);
    const uint32_t o = ldg2(P.nos + e.w);
    P.gates[g] = make_uint4(e.x, ldg2(P.nos + e.y), ldg2(P.nos + e.z), o);
    atomicMax(P.prod1 + o, g + 1);
  }
  grid_bar(P, cx);
  // ================= F8: deps (K2) =================
  {
    uint32_t f = 0;
    const uint32_t Gw = (G + 31) & ~31u;
    FUSED_FOR(g, Gw) {
      bool fwd = false;
      if (g < G) {
        const uint4 gt = P.gates[g];  // written by this very thread in F7 (same grid-stride mapping)
        const uint32_t d0 = ldg2(P.prod1 + gt.y) - 1u, d1 = ldg2(P.prod1 + gt.z) - 1u;
        P.dep[g] = make_uint2(d0, d1);
        fwd = (d0 != kNone && d0 > g) || (d1 != kNone && d1 > g);
        if (fwd) f |= F_OOO;
        if (d0 == g || d1 == g) f |= F_SELF;
      }
      warp_append(fwd, g, P.heavy, sc + FS_SEEDN);  // the seeds live in heavy[] until the roots phase reuses it
    }
    f = warp_or(f);
    if (lane == 0 && f) atomicOr(sc + FS_BFLAGS, f);
  }
  grid_bar(P, cx);
  const uint32_t bflags = ldg2(sc + FS_BFLAGS);
  const bool sorted = (bflags & (F_OOO | F_SELF)) != 0;  // otherwise the DFS post-order is 0..G-1
  if (sorted) {
    // ================= F9: exact DFS order (K5a-c) =================
    const uint32_t ns = ldg2(sc + FS_SEEDN);
    const bool sizable = G > 4096;  // below that even the quadratic worst case is microseconds
    bool deep = sizable && ns > G / 4;  // a forward edge in every fourth gate: the jumping stage of the multi-kernel path is the tool
    if (!deep) {  // seed: r[d] can only be lowered along a forward edge
      FUSED_FOR(i, ns) { uint32_t u = ldg2(P.heavy + i); fused_relax_from(u, u, P.dep, P.r, P.inq, P.q0, sc + FS_QN, kSeedHopCap, sc + FS_NCUT); }
      grid_bar(P, cx);
      uint32_t slot = 0, rr = 0, prev_nq = 0xFFFFFFFFu;
      uint32_t* qin = P.q0;
      uint32_t* qout = P.q1;
      while (true) {
        const uint32_t nq = ldg2(sc + FS_QN + slot), ncut = ldg2(sc + FS_NCUT);
        if (!nq) break;
        // the criteria of k_relax_loop: walks keep running into the cap, or the queue stays long (one hop per round through rh)
        if (sizable && ((rr == 0 && ncut > G / 8) || (rr >= 1 && ncut > G / 64) || (rr >= 3 && nq > G / 64 && (unsigned long long)nq * 8 > (unsigned long long)prev_nq * 7) ||
                        (rr >= kPatientRounds && ncut) || rr >= kMaxDataRounds)) { deep = true; break; }
        prev_nq = nq;
        const uint32_t nslot = (slot + 1) & 3;
        grid_bar(P, cx);  // everybody has read the cut count of the previous phase
        if (blockIdx.x == 0 && threadIdx.x == 0) { sc[FS_QN + ((nslot + 1) & 3)] = 0; sc[FS_NCUT] = 0; }
        grid_bar(P, cx);
        FUSED_FOR(i, nq) {
          uint32_t x = ldg2(qin + i);
          atomicAnd(P.inq + (x >> 5), ~(1u << (x & 31)));
          __threadfence();
          fused_relax_from(x, __ldcg(P.r + x), P.dep, P.r, P.inq, qout, sc + FS_QN + nslot, kRoundHopCap, sc + FS_NCUT);
        }
        grid_bar(P, cx);
        uint32_t* t = qin; qin = qout; qout = t;
        slot = nslot;
        ++rr;
      }
      if (blockIdx.x == 0 && threadIdx.x == 0) sc[FS_RELAX_ROUNDS] = rr;
    }
    if (deep) {  // uniform: every CTA derived it from the same scalars
      if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(sc + FS_EFLAGS, EF_DEEP);
      fused_publish(P);
      return;
    }
    FUSED_FOR(v, G) atomicAdd(P.size_off + ldg2(P.r + v), 1u);
    grid_bar(P, cx);
    {
      auto f = [&](uint32_t i) { return ldg2(P.size_off + i); };
      uint32_t tot = grid_scan(P, cx, G, f, P.size_off, P.agg + gridDim.x, s_warp);
      if (blockIdx.x == 0 && threadIdx.x == 0) P.size_off[G] = tot;
    }
    grid_bar(P, cx);
    {
      const bool check_self = bflags & F_SELF;
      FUSED_FOR(v, G) {
        if (ldg2(P.r + v) != v) continue;
        uint32_t o = ldg2(P.size_off + v), sz = ldg2(P.size_off + v + 1) - o;
        if (sz == 1) {
          if (check_self) {
            uint2 d = ldg2(P.dep + v);
            if (d.x == v || d.y == v) atomicMax(reinterpret_cast<unsigned long long*>(sc + FS_ERR_LO), ~(((unsigned long long)v << 32) | v));
          }
          order[o] = v;
        } else {
          P.heavy[atomicAdd(sc + FS_HEAVYN, 1u)] = v;
          if (sz >= 4 * kBigBlock) atomicOr(sc + FS_EFLAGS, EF_DEEP);  // a one-thread DFS of thousands of gates: k_tree_blocks' job
        }
      }
    }
    grid_bar(P, cx);
    if (ldg2(sc + FS_EFLAGS) & EF_DEEP) { fused_publish(P); return; }
    {
      const uint32_t nh = ldg2(sc + FS_HEAVYN);
      FUSED_FOR(i, nh) {
        const uint32_t R = ldg2(P.heavy + i);
        const uint32_t base = ldg2(P.size_off + R), end = ldg2(P.size_off + R + 1);
        uint32_t emit = base, top = end;
        P.state[R] = 1;
        order[--top] = R;
        while (top < end) {
          uint32_t v = order[top];
          uint8_t s = P.state[v];
          if (s <= 2) {
            uint2 dd = ldg2(P.dep + v);
            uint32_t d = (s == 1) ? dd.x : dd.y;
            P.state[v] = s + 1;
            if (d != kNone && ldg2(P.r + d) == R) {
              uint8_t sd = P.state[d];
              if (sd == 0) { P.state[d] = 1; order[--top] = d; }
              else if (sd < 4) {  // visiting[d]  (topological_sort.rs:34-38)
                atomicMax(reinterpret_cast<unsigned long long*>(sc + FS_ERR_LO), ~(((unsigned long long)R << 32) | d));
                break;
              }
            }
          } else { ++top; order[emit++] = v; P.state[v] = 4; }
        }
      }
    }
    grid_bar(P, cx);
    if (ldg2(sc + FS_ERR_HI) | ldg2(sc + FS_ERR_LO)) { fused_publish(P); return; }  // cyclic dependency: the host reports "detected at i="
  } else {
    FUSED_FOR(i, G) order[i] = i;
  }
  // ================= F10: wire numbering (K6) + gather (K7) =================
  FUSED_FOR(k, G) {
    const uint32_t g = sorted ? ldg2(order + k) : k;
    const uint4 gt = ldg2(P.gates + g);
    const uint2 dd = ldg2(P.dep + g);
    const uint32_t p = kFirstTag | (3u * k);
    // an operand whose node has a producer cannot hold the node's first appearance (see k_wire_first)
    if (dd.x == kNone) fused_red_min(wire + gt.y, p);
    if (dd.y == kNone && gt.z != gt.y) fused_red_min(wire + gt.z, p + 1);
    if (gt.w != gt.y && gt.w != gt.z) fused_red_min(wire + gt.w, p + 2);
  }
  grid_bar(P, cx);
  FUSED_FOR(nd, NB) {
    const uint32_t w = ldg2(wire + nd);
    if ((w & kFirstTag) && w != kNone) { uint32_t p = w & ~kFirstTag; atomicOr(P.bitmap + (p >> 5), 1u << (p & 31)); }
  }
  grid_bar(P, cx);
  uint32_t n_mid;
  {
    const uint32_t BW = (3 * G + 31) / 32 + 1;
    auto f = [&](uint32_t w) { return (uint32_t)__popc(ldg2(P.bitmap + w)); };
    n_mid = grid_scan(P, cx, BW, f, P.bitmap_pre, P.agg + 2 * gridDim.x, s_warp);
    if (blockIdx.x == 0 && threadIdx.x == 0) sc[FS_NMID] = n_mid;
  }
  grid_bar(P, cx);
  FUSED_FOR(nd, NB) {
    const uint32_t w = ldg2(wire + nd);
    if (w == kNone) continue;
    uint32_t id;
    if (w & kFirstTag) {
      const uint32_t p = w & ~kFirstTag;
      id = P.n_in + ldg2(P.bitmap_pre + (p >> 5)) + __popc(ldg2(P.bitmap + (p >> 5)) & ((1u << (p & 31)) - 1u));
    } else if (w > kOutBase) id = kInBase - w;               // input: its list position
    else id = P.n_in + n_mid + (kOutBase - w);                // output: after all intermediates
    wire[nd] = id;
  }
  grid_bar(P, cx);
  if (new_gates) {
    FUSED_FOR(k, G) {
      const uint32_t g = sorted ? ldg2(order + k) : k;
      const uint4 gt = ldg2(P.gates + g);
      new_gates[k] = make_uint4(gt.x, ldg2(wire + gt.y), ldg2(wire + gt.z), ldg2(wire + gt.w));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc[FS_DONE] = 1u;
    if (P.trace) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); P.trace[126] = t; }
  }
  fused_publish(P);
}

}  // namespace c2a

using namespace c2a;

extern "C" {

// 0 = never take the fused kernel; otherwise the largest event count it is used for (default: its structural limit)
static uint64_t g_fused_max_events = (uint64_t)kFusedMaxTiles * kEvTile;
static uint32_t g_fused_events_per_cta = 1024;  // one event per thread and pass: the phases are latency-bound (measured: 4096 is 10-25 % slower below 100 K gates)
void c2a_set_fused_limits(uint64_t max_events, uint32_t events_per_cta) {
  g_fused_max_events = std::min<uint64_t>(max_events, (uint64_t)kFusedMaxTiles * kEvTile);
  if (events_per_cta) g_fused_events_per_cta = events_per_cta;
}

static int compile_packed_impl(c2a_handle* h, const c2a_packed_events* pk, bool pk_on_device, const c2a_compile_io* io, bool out_on_device,
                               c2a_emit_info* info, uint32_t* wire_count, uint64_t* err_event, uint64_t* err_index) {
  if (!h) return C2A_ERR_INVALID_ARGUMENT;
  if (!pk || !io) return fail(h, C2A_ERR_INVALID_ARGUMENT, "null argument");
  const uint64_t n = pk->n_events, nw = pk->n_words;
  // the multi-kernel pipeline: any size, any stream, exact error replay.  defer_ok: the emit's final status is read together with the
  // build's (one synchronisation less); a stream that fails the deferred check is replayed with defer_ok = false.
  std::function<int(bool)> classic_run = [&](bool defer_ok) -> int {
    EmitSrc src;
    src.pk = pk;
    src.pk_on_device = pk_on_device;
    EmitDefer df;
    df.wire_cap = io->wire_of_node ? io->wire_cap : 0u;  // host array or device array: nothing may be written behind wire_cap
    // the I/O signal lists go up on the side stream before anything else (the emit's scatter joins that stream: they have landed
    // long before the build maps them to nodes)
    const uint64_t n_io_all = (uint64_t)io->n_in + io->n_out;
    const uint32_t* d_io_sigs = nullptr;
    if (defer_ok && n_io_all && n_io_all <= (1u << 24) && (pk->flags & C2A_PACKED_DENSE_IDS)) {
      if ((4 * n_io_all + 2048) > h->h_pinned_bytes) {
        cudaStreamSynchronize(h->stream);
        if (h->h_pinned) cudaFreeHost(h->h_pinned);
        h->h_pinned_bytes = 4 * n_io_all + 8192;
        if (!cuda_ok(h, cudaHostAlloc((void**)&h->h_pinned, h->h_pinned_bytes, cudaHostAllocDefault), "cudaHostAlloc")) return C2A_ERR_CUDA;
      }
      if (4 * n_io_all + 16 > h->io_bytes) {
        if (h->io_buf) { cudaStreamSynchronize(h->stream2); cudaFree(h->io_buf); h->io_buf = nullptr; h->io_bytes = 0; }
        if (cudaMalloc(&h->io_buf, 4 * n_io_all + 4096) == cudaSuccess) h->io_bytes = 4 * n_io_all + 4096;
        else cudaGetLastError();
      }
      if (h->io_buf) {
        cudaEventRecord(h->ev_main, h->stream);  // (behind anything an earlier, failed call left running: it may still read io_buf)
        cudaStreamWaitEvent(h->stream2, h->ev_main, 0);
        uint32_t* stage = h->h_pinned + 256;
        if (io->n_in) memcpy(stage, io->input_signals, 4 * (size_t)io->n_in);
        if (io->n_out) memcpy(stage + io->n_in, io->output_signals, 4 * (size_t)io->n_out);
        cudaMemcpyAsync(h->io_buf, stage, 4 * n_io_all, cudaMemcpyHostToDevice, h->stream2);
        d_io_sigs = (const uint32_t*)h->io_buf;
      }
    }
    h->phase_prefix = "emit:";
    int st = emit_events_impl(h, src, n, info, err_event, defer_ok ? &df : nullptr);
    h->phase_prefix.clear();
    if (st != C2A_OK) return st;
    // a deferred emit has not synchronised: every way out of this function below does (the build's final synchronisation, or here)
    auto quiesce = [&]() { if (df.pending) { cudaStreamSynchronize(h->stream2); cudaStreamSynchronize(h->stream); } };
    if ((io->order_out || io->new_gates) && io->gates_cap < h->emitted.G) {
      if (df.pending) {  // report it the undeferred way: *info complete, and a stream the reference rejects keeps its own error
        quiesce();
        slab_reset(h);
        return classic_run(false);
      }
      return fail(h, C2A_ERR_INVALID_ARGUMENT, "gates_cap (%llu) < number of gates (%llu)", (unsigned long long)io->gates_cap, (unsigned long long)h->emitted.G); }
    if (!df.pending && io->wire_of_node && io->wire_cap < h->emitted.node_count + 1) return fail(h, C2A_ERR_INVALID_ARGUMENT, "wire_cap (%u) < node_count + 1 (%u)", io->wire_cap, h->emitted.node_count + 1);
    st = emitted_build_impl(h, io->input_signals, io->n_in, io->output_signals, io->n_out, io->order_out, io->wire_of_node, io->new_gates, wire_count, err_index, out_on_device,
                            0, ~0ull, /*keep_phases=*/true, df.pending ? d_io_sigs : nullptr, df.pending ? df.es : nullptr);
    if (!df.pending) return st;
    if (st != C2A_OK) quiesce();  // (an error from before the build's own synchronisation: sizes, memory)
    // ---- the emit's own status, now that the stream has been synchronised
    const uint32_t* es = h->h_emit_status;
    uint32_t flags = es[ES_FLAGS];
    if (es[ES_NDECL] != df.n_sig) flags |= EF_DUPLICATE;
    const bool converged = !(df.C && es[ES_MC0 + 2 * (kSpecMsf - 1)] != 0 && es[ES_MC0 + 2 * (kSpecMsf - 1) + 1] != 0);
    if (flags || !converged) {  // the reference rejects the stream, or the forest needed more Boruvka rounds: the ordinary way
      slab_reset(h);
      return classic_run(false);
    }
    const uint32_t n_eff = df.C ? es[ES_COUNT] : 0u;
    const uint32_t node_count = (uint32_t)(df.n_sig + n_eff);
    h->emitted.node_count = node_count;
    if (info) { info->n_effective = n_eff; info->node_count = node_count; info->path = C2A_EMIT_PATH_DEVICE; info->rounds = es[ES_ROUNDS]; }
    if (io->wire_of_node && io->wire_cap < node_count + 1) return fail(h, C2A_ERR_INVALID_ARGUMENT, "wire_cap (%u) < node_count + 1 (%u)", io->wire_cap, node_count + 1);
    return st;
  };
  // (a wire map without room for a single entry can only fail: the undeferred form reports it before anything is built)
  auto classic = [&]() -> int { return classic_run(!(io->wire_of_node && io->wire_cap == 0)); };
  const bool dense = (pk->flags & C2A_PACKED_DENSE_IDS) != 0;
  const uint64_t n_io = (uint64_t)io->n_in + io->n_out;
  if (!dense || n == 0 || n > g_fused_max_events || nw > 3 * n || n_io > (1u << 24) || (n && !pk->kinds) || (nw && !pk->words)) return classic();
  int st = check_sizes(h, n, 1);
  if (st) return st;
  if (info) { memset(info, 0, sizeof *info); info->n_events = n; }
  phases_clear(h);
  cudaStream_t s = h->stream;

  FusedParams P;
  memset(&P, 0, sizeof P);
  P.n = (uint32_t)n;
  P.n_words = (uint32_t)nw;
  P.tiles = (uint32_t)((n + kEvTile - 1) / kEvTile);
  P.n_in = io->n_in;
  P.n_out = io->n_out;
  const bool implicit = (pk->flags & C2A_PACKED_IMPLICIT_OPERANDS) != 0;  // a gate carries >= 2 words, a connection >= 1
  P.implicit = implicit ? 1u : 0u;
  P.G_ub = (uint32_t)(implicit ? nw / 2 : nw / 3);
  P.C_ub = (uint32_t)(implicit ? nw : nw / 2);
  P.S_ub = (uint32_t)(n - (nw + 2) / 3 + 1);
  P.NB_ub = P.S_ub + P.C_ub + 1;
  const uint64_t Gu = P.G_ub, Cu = P.C_ub, Su = P.S_ub, NBu = P.NB_ub;
  int grid = (int)std::min<uint64_t>((uint64_t)h->num_sms, std::max<uint64_t>(1, (n + g_fused_events_per_cta - 1) / g_fused_events_per_cta));
  {
    static int occ = -1;
    if (occ < 0) {
      if (cudaFuncSetAttribute((const void*)k_fused_compile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem) != cudaSuccess ||
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k_fused_compile, kFusedBlock, kFusedSmem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 0; }
    }
    if (occ < 1) return classic();
  }
  // ---- staging for the stream / I/O lists (host form), then one slab: resident results first, scratch behind them
  const size_t kbytes = align256(n + 16), wbytes = align256(4 * nw + 16), iobytes = align256(4 * n_io + 16);
  const size_t stage_need = (pk_on_device ? 0 : kbytes + wbytes) + iobytes;
  if (stage_need > h->ev_bytes) {
    if (h->ev_buf) { cudaStreamSynchronize(s); cudaFree(h->ev_buf); h->ev_buf = nullptr; h->ev_bytes = 0; }
    if (!cuda_ok(h, cudaMalloc(&h->ev_buf, stage_need + stage_need / 8 + 4096), "cudaMalloc(event staging)")) { cudaGetLastError(); return C2A_ERR_NO_MEMORY; }
    h->ev_bytes = stage_need + stage_need / 8 + 4096;
  }
  if ((4 * n_io + 8192) > h->h_pinned_bytes) {
    cudaStreamSynchronize(s);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    h->h_pinned_bytes = 4 * n_io + 32768;
    if (!cuda_ok(h, cudaHostAlloc((void**)&h->h_pinned, h->h_pinned_bytes, cudaHostAllocDefault), "cudaHostAlloc")) return C2A_ERR_CUDA;
  }
  char* stg = h->ev_buf;
  uint32_t* d_io = (uint32_t*)stg;
  stg += iobytes;
  if (pk_on_device) { P.kinds = pk->kinds; P.words = pk->words; }
  else { P.kinds = (const uint8_t*)stg; P.words = (const uint32_t*)(stg + kbytes); }
  P.io_sigs = d_io;

  slab_reset(h);
  emit_drop_host(h);
  size_t need = align256(16 * Gu) + align256(4 * Su) + align256(4 * NBu);                                            // resident
  need += 3 * align256(4 * ((size_t)P.tiles + 2)) + align256(8 * Su) + align256(16 * Gu) + align256(8 * Cu) + align256(4 * Cu) + align256(Su);  // scatter targets
  need += 5 * align256(4 * Su) + 2 * align256(4 * (Cu / 32 + 8)) + align256(4 * Cu) + align256(16 * Cu);            // union-find, bitmaps, live lists
  need += align256(8 * Gu) + 2 * align256(4 * (Gu + 1)) + align256(Gu) + align256(4 * (Gu / 32 + 2)) + 3 * align256(4 * Gu) + 2 * align256(4 * ((3 * Gu + 31) / 32 + 8));
  need += align256(4 * Gu) + align256(4 * NBu) + align256(16 * Gu);                                                 // order / wire / new gates when not the caller's
  need += align256(4 * FS_COUNT + 64 + 8 * 3 * (size_t)grid) + 4096;
  if (!slab_reserve(h, need)) return C2A_ERR_NO_MEMORY;
  auto A = [&](size_t bytes) { return slab_alloc(h, bytes); };
  P.gates = (uint4*)A(16 * Gu);
  P.nos = (uint32_t*)A(4 * Su);
  P.prod1 = (uint32_t*)A(4 * NBu);
  const size_t keep = h->slab_used;
  P.tile_g = (uint32_t*)A(4 * ((size_t)P.tiles + 2));
  P.tile_c = (uint32_t*)A(4 * ((size_t)P.tiles + 2));
  P.tile_i = (uint32_t*)A(4 * ((size_t)P.tiles + 2));
  P.sig_meta = (uint2*)A(8 * Su);
  P.egates = (uint4*)A(16 * Gu);
  P.conn = (uint2*)A(8 * Cu);
  P.conn_sb = (uint32_t*)A(4 * Cu);
  P.outmark = (uint8_t*)A(Su);
  P.parent = (uint32_t*)A(4 * Su);
  P.best = (uint32_t*)A(4 * Su);
  P.nidf = (uint32_t*)A(4 * Su);
  P.in_idx1 = (uint32_t*)A(4 * Su);
  P.out_idx1 = (uint32_t*)A(4 * Su);
  P.eff = (uint32_t*)A(4 * (Cu / 32 + 8));
  P.effp = (uint32_t*)A(4 * (Cu / 32 + 8));
  P.cur = (uint32_t*)A(4 * Cu);
  P.cand = (uint4*)A(16 * Cu);
  P.dep = (uint2*)A(8 * Gu);
  P.r = (uint32_t*)A(4 * (Gu + 1));
  P.size_off = (uint32_t*)A(4 * (Gu + 1));
  P.state = (uint8_t*)A(Gu);
  P.inq = (uint32_t*)A(4 * (Gu / 32 + 2));
  P.q0 = (uint32_t*)A(4 * Gu);
  P.q1 = (uint32_t*)A(4 * Gu);
  P.heavy = (uint32_t*)A(4 * Gu);
  P.bitmap = (uint32_t*)A(4 * ((3 * Gu + 31) / 32 + 8));
  P.bitmap_pre = (uint32_t*)A(4 * ((3 * Gu + 31) / 32 + 8));
  // results: the caller's device arrays are handed to the kernel together with their capacities (it writes them itself once the
  // exact sizes are known to fit); internal arrays of bound size stand by for host destinations and for capacities that do not fit
  const bool want_new = io->new_gates != nullptr;
  const bool user_dev = out_on_device;
  const bool gates_may_fit = user_dev && io->gates_cap >= 1;
  P.order_user = (user_dev && io->order_out) ? io->order_out : nullptr;
  P.new_user = (user_dev && want_new) ? (uint4*)io->new_gates : nullptr;
  P.wire_user = (user_dev && io->wire_of_node) ? io->wire_of_node : nullptr;
  P.gates_cap = io->gates_cap;
  P.wire_cap = io->wire_cap;
  // (a device caller whose capacities cover the bounds never needs the internal copies; otherwise they are carved - cheap, the slab
  //  is reused - so that a capacity that turns out too small cannot make the kernel write out of bounds)
  P.order_int = (P.order_user && io->gates_cap >= Gu) ? P.order_user : (uint32_t*)A(4 * Gu);
  P.wire_int = (P.wire_user && io->wire_cap >= NBu) ? P.wire_user : (uint32_t*)A(4 * NBu);
  P.new_int = !want_new ? nullptr : ((P.new_user && io->gates_cap >= Gu) ? P.new_user : (uint4*)A(16 * Gu));
  (void)gates_may_fit;
  static const bool want_trace = getenv("C2A_FUSED_TRACE") != nullptr;
  unsigned long long* d_trace = want_trace ? (unsigned long long*)A(1024) : nullptr;
  if (!P.order_int || !P.wire_int || (want_new && !P.new_int) || (want_trace && !d_trace)) return fail(h, C2A_ERR_NO_MEMORY, "scratch slab exhausted");
  // control block: two halves in the handle, zeroed once; every launch zeroes the half the NEXT launch uses
  const size_t ctl_half = align256(4 * FS_COUNT + 64 + 8 * 3 * (size_t)h->num_sms);
  if (!h->fused_ctl) {
    if (!cuda_ok(h, cudaMalloc(&h->fused_ctl, 2 * ctl_half), "cudaMalloc(fused control)")) { cudaGetLastError(); return C2A_ERR_NO_MEMORY; }
    cudaMemsetAsync(h->fused_ctl, 0, 2 * ctl_half, s);
    h->fused_parity = 0;
  }
  char* ctl = h->fused_ctl + (h->fused_parity ? ctl_half : 0);
  P.sc = (uint32_t*)ctl;
  P.bar = (unsigned int*)(ctl + 4 * FS_COUNT);
  P.agg = (unsigned long long*)(ctl + 4 * FS_COUNT + 64);
  P.ctl_next = (uint32_t*)(h->fused_ctl + (h->fused_parity ? 0 : ctl_half));
  P.ctl_words = (uint32_t)(ctl_half / 4);
  P.trace = d_trace;
  h->fused_parity ^= 1;

  // ---- enqueue: (copies in,) the kernel, (copies out); ONE synchronisation.  Scalars come back through mapped pinned memory,
  // short I/O lists are read by the kernel from pinned memory directly.
  uint32_t* hp = h->h_pinned;
  P.host_sc = hp;
  memset(hp, 0, 4 * FS_COUNT);
  if (n_io) {
    uint32_t* stage = hp + 512;
    if (io->n_in) memcpy(stage, io->input_signals, 4 * (size_t)io->n_in);
    if (io->n_out) memcpy(stage + io->n_in, io->output_signals, 4 * (size_t)io->n_out);
    if (n_io <= 4096) P.io_sigs = stage;  // zero-copy: a few KB over PCIe inside F2
    else cudaMemcpyAsync(d_io, stage, 4 * n_io, cudaMemcpyHostToDevice, s);
  }
  if (!pk_on_device) {
    phase_begin(h, "h2d");
    if (!cuda_ok(h, cudaMemcpyAsync((void*)P.kinds, pk->kinds, n, cudaMemcpyHostToDevice, s), "kinds H2D")) return C2A_ERR_CUDA;
    if (nw && !cuda_ok(h, cudaMemcpyAsync((void*)P.words, pk->words, 4 * nw, cudaMemcpyHostToDevice, s), "words H2D")) return C2A_ERR_CUDA;
    phase_end(h);
  }
  if (want_trace) cudaMemsetAsync(d_trace, 0, 1024, s);
  phase_begin(h, "k_fused_compile");
  {
    void* args[] = {(void*)&P};
    static const bool no_cluster = getenv("C2A_FUSED_NO_CLUSTER") != nullptr;
    bool launched = false;
    if (grid <= 8 && !no_cluster) {  // one cluster: co-scheduled by construction, hardware barrier
      P.cluster = 1;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(kFusedBlock);
      cfg.dynamicSmemBytes = kFusedSmem;
      cfg.stream = s;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = grid;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      if (cudaLaunchKernelExC(&cfg, (const void*)k_fused_compile, args) == cudaSuccess) launched = true;
      else { cudaGetLastError(); P.cluster = 0; }  // (no room for the cluster right now: the cooperative form below)
    }
    if (!launched && !cuda_ok(h, cudaLaunchCooperativeKernel((const void*)k_fused_compile, dim3(grid), dim3(kFusedBlock), args, kFusedSmem, s), "fused launch")) return C2A_ERR_CUDA;
    h->launches++;
  }
  phase_end(h);
  if (want_trace) cudaMemcpyAsync(hp + 64, d_trace, 1024, cudaMemcpyDeviceToHost, s);
  // host destinations: copied by capacity (the exact sizes are only known after the synchronisation; a caller that sized its arrays
  // exactly copies exactly)
  if (!out_on_device) {
    const uint64_t gcopy = std::min<uint64_t>(io->gates_cap, Gu), wcopy = std::min<uint64_t>(io->wire_cap, NBu);
    phase_begin(h, "d2h");
    if (io->order_out && gcopy) cudaMemcpyAsync(io->order_out, P.order_int, 4 * gcopy, cudaMemcpyDeviceToHost, s);
    if (io->wire_of_node && wcopy) cudaMemcpyAsync(io->wire_of_node, P.wire_int, 4 * wcopy, cudaMemcpyDeviceToHost, s);
    if (want_new && gcopy) cudaMemcpyAsync(io->new_gates, P.new_int, 16 * gcopy, cudaMemcpyDeviceToHost, s);
    phase_end(h);
  }
  if (!cuda_ok(h, cudaStreamSynchronize(s), "fused sync")) return C2A_ERR_CUDA;
  if (!cuda_ok(h, cudaGetLastError(), "fused kernel")) return C2A_ERR_CUDA;
  phases_collect(h);
  if (want_trace) {  // per-barrier timeline of CTA 0 as extra "phases": fused:b<k> = ms between barrier k-1 and k
    const unsigned long long* tr = (const unsigned long long*)(hp + 64);
    unsigned long long prev = tr[127];
    for (unsigned k = 1; k <= tr[0] && k < 120; ++k) {
      char nm[32];
      const unsigned long long t = tr[k] & ~(1ull << 63);
      snprintf(nm, sizeof nm, (tr[k] >> 63) ? "fused:m%02u" : "fused:b%02u", k);
      h->last_ms.push_back({nm, (double)(t - prev) * 1e-6});
      prev = t;
    }
    if (tr[126]) h->last_ms.push_back({"fused:tail", (double)(tr[126] - prev) * 1e-6});
    h->last_ms.push_back({"fused:grid", (double)grid});
  }

  const uint32_t G = hp[FS_G], Cn = hp[FS_C], S = hp[FS_S], eflags = hp[FS_EFLAGS];
  const unsigned long long err_enc = ((unsigned long long)hp[FS_ERR_HI] << 32) | hp[FS_ERR_LO];
  if (eflags || (!hp[FS_DONE] && !err_enc)) {
    // the reference errors on this stream, or it is not a stream the device path decides itself: the multi-kernel path replays it
    // exactly (same status, same event index).  Nothing of the fused attempt is kept.
    slab_reset(h);
    return classic();
  }
  const uint32_t n_eff = hp[FS_NEFF];
  if (info) {
    info->n_gates = G; info->n_connections = Cn; info->n_signals = S; info->signal_bound = S; info->n_effective = n_eff;
    info->node_count = S + n_eff; info->path = C2A_EMIT_PATH_DEVICE; info->rounds = hp[FS_ROUNDS];
  }
  h->slab_keep = keep;
  h->emitted.valid = true;
  h->emitted.nos_valid = true;
  h->emitted.gates_off = (char*)P.gates - h->slab;
  h->emitted.nos_off = (char*)P.nos - h->slab;
  h->emitted.prod1_valid = true;
  h->emitted.prod1_off = (char*)P.prod1 - h->slab;
  h->emitted.G = G;
  h->emitted.node_count = S + n_eff;
  h->emitted.signal_bound = S;
  h->emitted.wire = nullptr;
  if (err_enc) {
    const unsigned long long err = ~err_enc;
    if (err_index) *err_index = (uint32_t)err;
    return fail(h, C2A_ERR_CYCLIC_DEPENDENCY, "detected at i=%llu", (unsigned long long)(uint32_t)err);
  }
  // which wire map the kernel filled (the same rule it applied): needed for the named-wire look-ups that may follow
  uint32_t* d_wire = (P.wire_user && S + n_eff + 1 <= io->wire_cap) ? P.wire_user : P.wire_int;
  h->emitted.wire = d_wire;
  h->emitted.identity = !(hp[FS_BFLAGS] & (F_OOO | F_SELF));
  if (wire_count) *wire_count = io->n_in + hp[FS_NMID] + io->n_out;
  if ((io->order_out || io->new_gates) && io->gates_cap < G) return fail(h, C2A_ERR_INVALID_ARGUMENT, "gates_cap (%llu) < number of gates (%u)", (unsigned long long)io->gates_cap, G);
  if (io->wire_of_node && io->wire_cap < S + n_eff + 1) return fail(h, C2A_ERR_INVALID_ARGUMENT, "wire_cap (%u) < node_count + 1 (%u)", io->wire_cap, S + n_eff + 1);
  return C2A_OK;
}

int c2a_compile_packed(c2a_handle* h, const c2a_packed_events* pk, const c2a_compile_io* io, c2a_emit_info* info, uint32_t* wire_count,
                       uint64_t* err_event, uint64_t* err_index) {
  return compile_packed_impl(h, pk, false, io, false, info, wire_count, err_event, err_index);
}
int c2a_compile_packed_resident(c2a_handle* h, const c2a_packed_events* d_pk, const c2a_compile_io* io, c2a_emit_info* info, uint32_t* wire_count,
                                uint64_t* err_event, uint64_t* err_index) {
  return compile_packed_impl(h, d_pk, true, io, true, info, wire_count, err_event, err_index);
}

}  // extern "C"
