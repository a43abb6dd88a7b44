// c2a_emit.cuh — device emitter: the reference's Compiler::{add_signal, add_gate, add_connection}
// (src/compiler.rs:139-278) replayed over a whole event stream ON THE GPU, producing the same node ids and the
// same node-id gate vector as the sequential reference, then feeding build_circuit without a host round trip.
// Included by c2a_device.cu (single translation unit).
//
// Why this is parallel at all.  The reference processes connections in event order; a connection is *effective*
// (consumes a node id, compiler.rs:257) iff its two signals are not already in one node (:235-237).  Processing
// edges in time order and keeping those that join two classes is Kruskal's algorithm with weight = event index,
// so the effective connections are exactly the edges of the (unique, weights are distinct) minimum spanning
// forest of the connection graph -> Boruvka rounds on the device.  Node ids then follow from two prefix counts:
//   * add_signal at event t:            id = #signals declared up to t  + #effective connections before t   (:157)
//   * effective connection at event t:  id = #signals declared before t + #effective connections up to t    (:257)
//   * a class ends with the id of its LAST event, i.e. the largest id among its members / MSF edges.
// gates hold node ids that the reference keeps current on every merge (:260-270): final value = final class id.
//
// What the device path does not decide itself: streams that make the reference return an error or take the
// "signal in no node => node 0" path (:183).  They are detected (flags below) and replayed by the exact host
// emitter (c2a_host.cpp), which yields the reference's error code / event index.  Detection is exact for
// duplicates, unknown references and const+const merges; out+out is flagged conservatively (a final class
// holding two distinct gate-output signals), which is exact for walker-generated streams where every gate
// writes a fresh temporary (src/process.rs:470-475).
//
//   E0 k_ev_count      per-1024-event tile: #gates, #connections; 1+max signal id; kind/op validation
//      k_scan_u32 x2   exclusive scans of the two per-tile count arrays (single-pass, decoupled look-back)
//   E1 k_ev_scatter    ballot/popc ranks inside a tile -> gate index / connection index / signal index of every
//                      event; scatters into SoA arrays (sig_t, sig_meta | egates, gate_t | conn, conn_t, conn_sb).
//                      E0..E1 run back to back: their targets are sized by upper bounds and live in the staging allocation,
//                      so the counts are only read by the host once, after the scatter.
//   E2 k_ev_check_*    every reference precedes its use (else node-0 semantics), marks gate-output signals
//   M  k_msf_pick / k_msf_hook   Boruvka: per class the minimum-time outgoing connection (tagged RED.MIN),
//                      hook, path compression inside find; edge list shrinks every round.  kSpecMsf rounds are enqueued
//                      unconditionally (an exhausted round exits at once); more only if the final status says so.
//   N  k_scan_u32<popc>(effective-connection bitmap), k_ev_nid_edges, k_ev_finalize, k_ev_gates
#pragma once

struct c2a_compiler;

namespace c2a {

constexpr int kEvTile = 1024;  // events per CTA pass: 8 warps x 4 rows x 32 lanes
constexpr int kSpecMsf = 2;    // Boruvka rounds enqueued without looking at the live-edge count
// ES_MC0 + 2r / + 2r + 1: candidate / live counts of speculative round r (zeroed once); ES_NCUR / ES_NCAND: the host-driven rounds
enum { ES_NGATE = 0, ES_NCONN = 1, ES_SBOUND = 2, ES_FLAGS = 3, ES_NCUR = 4, ES_NCAND = 5, ES_NDECL = 6, ES_PREV = 7, ES_ROUNDS = 8, ES_IOBAD = 9, ES_NIMPL = 10,
       ES_MC0 = 16, ES_COUNT = 32 };
enum { EF_BAD_KIND = 1, EF_BAD_OP = 2, EF_DUPLICATE = 4, EF_UNKNOWN_REF = 8, EF_CONST_CONST = 16, EF_OUT_OUT = 32, EF_SPARSE = 64, EF_BAD_IO = 128,
       EF_CAP = 256 /* a signal id beyond the table bound the pass was launched with: rerun with the exact bound */ };

// ---- E0 ---------------------------------------------------------------------------------------------------
// (A single fused pass - tile counts chained by a decoupled look-back inside the scatter - was measured at 0.67 ms on the
//  42 M-event stream: every tile then carries ticket + load + look-back latency in series.  Reading the stream twice is faster.)
__global__ void __launch_bounds__(kBlock) k_ev_count(const uint4* __restrict__ ev, uint64_t n, uint32_t tiles, uint32_t* __restrict__ tile_g,
                                                     uint32_t* __restrict__ tile_c, uint32_t* __restrict__ es) {
  __shared__ uint32_t s_g[8], s_c[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t smax = 0, f = 0;
  for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    uint64_t base = (uint64_t)tile * kEvTile + warp * 128;
    uint32_t g = 0, c = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint64_t i = base + j * 32 + lane;
      if (i < n) {
        uint4 e = ldg_stream(ev + i);
        uint32_t kind = e.x & 0xFF;
        if (kind <= C2A_EV_SIGNAL_CONST) {
          if (e.y == 0xFFFFFFFFu) f |= EF_SPARSE; else smax = max(smax, e.y + 1);
        } else if (kind == C2A_EV_GATE) {
          ++g;
          if ((e.x >> 8) >= C2A_GATE_TYPE_COUNT) f |= EF_BAD_OP;
        } else if (kind == C2A_EV_CONNECT) ++c;
        else f |= EF_BAD_KIND;
      }
    }
    g = warp_sum(g);
    c = warp_sum(c);
    if (lane == 0) { s_g[warp] = g; s_c[warp] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t tg = 0, tc = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) { tg += s_g[w]; tc += s_c[w]; }
      tile_g[tile] = tg;
      tile_c[tile] = tc;
    }
    __syncthreads();
  }
  smax = warp_max(smax);
  f = warp_or(f);
  if (lane == 0) {
    if (smax) atomicMax(es + ES_SBOUND, smax);
    if (f) atomicOr(es + ES_FLAGS, f);
  }
}

// ---- E1 ---------------------------------------------------------------------------------------------------
// tile_g / tile_c: exclusive scans of the per-tile counts (k_scan_u32), entry [tiles] = totals.
// sig_t[sid]    event index of the declaration (kNone = never declared)
// sig_meta[sid] {signal index (declaration rank) | is_const << 31, #connections before the declaration}
// (dense ids: one 4-byte word per signal instead, #connections before the declaration | is_const << 31)
// egates[g]     {op, lhs signal, rhs signal, out signal};  gate_t[g] event index
// conn[c]       {a, b};  conn_t[c] event index;  conn_sb[c] #signals declared before the connection
__global__ void __launch_bounds__(kBlock) k_ev_scatter(const uint4* __restrict__ ev, uint64_t n, uint32_t tiles, uint32_t S_cap,
                                                       const uint32_t* __restrict__ tile_g, const uint32_t* __restrict__ tile_c,
                                                       uint32_t* __restrict__ sig_t, uint2* __restrict__ sig_meta, uint4* __restrict__ egates,
                                                       uint32_t* __restrict__ gate_t, uint2* __restrict__ conn, uint32_t* __restrict__ conn_t,
                                                       uint32_t* __restrict__ conn_sb, uint32_t* __restrict__ es) {
  __shared__ uint32_t s_g[8], s_c[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t f = 0;
  for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    uint64_t base = (uint64_t)tile * kEvTile + warp * 128;
    uint4 e[4];
    uint32_t gm[4], cm[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint64_t i = base + j * 32 + lane;
      e[j] = make_uint4(0xFFu, 0, 0, 0);
      if (i < n) e[j] = ldg_stream(ev + i);
    }
    uint32_t gi = __ldg(tile_g + tile), ci = __ldg(tile_c + tile);
    uint32_t wg = 0, wc = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t kind = e[j].x & 0xFF;
      gm[j] = __ballot_sync(0xFFFFFFFFu, kind == C2A_EV_GATE);
      cm[j] = __ballot_sync(0xFFFFFFFFu, kind == C2A_EV_CONNECT);
      wg += __popc(gm[j]);
      wc += __popc(cm[j]);
    }
    if (lane == 0) { s_g[warp] = wg; s_c[warp] = wc; }
    __syncthreads();
    for (int w = 0; w < warp; ++w) { gi += s_g[w]; ci += s_c[w]; }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint64_t i = base + j * 32 + lane;
      uint32_t my_g = gi + __popc(gm[j] & lt), my_c = ci + __popc(cm[j] & lt);
      gi += __popc(gm[j]);
      ci += __popc(cm[j]);
      if (i >= n) continue;
      uint32_t kind = e[j].x & 0xFF;
      uint32_t t = (uint32_t)i;
      uint32_t my_s = t - my_g - my_c;  // signals declared before this event
      if (kind <= C2A_EV_SIGNAL_CONST) {
        // plain stores: a duplicate declaration (compiler.rs:146-148) overwrites, and is caught later because the
        // number of declared ids then falls short of the number of signal events (k_ev_finalize counts them)
        uint32_t sid = e[j].y;
        if (sid >= S_cap) f |= EF_CAP;
        else {
          sig_t[sid] = t;
          sig_meta[sid] = make_uint2(my_s | (kind == C2A_EV_SIGNAL_CONST ? 0x80000000u : 0u), my_c);
        }
      } else if (kind == C2A_EV_GATE) {
        // operands are range-checked (and neutralised) by k_ev_check_gates
        egates[my_g] = make_uint4(e[j].x >> 8, e[j].y, e[j].z, e[j].w);
        gate_t[my_g] = t;
      } else if (kind == C2A_EV_CONNECT) {
        conn[my_c] = make_uint2(e[j].y, e[j].z);
        conn_t[my_c] = t;
        conn_sb[my_c] = my_s;
      }
    }
  }
  f = warp_or(f);
  if (lane == 0 && f) atomicOr(es + ES_FLAGS, f);
}

// ---- E0 / E1 on a PACKED stream (include/c2a.h: kinds byte + payload words) ------------------------------------------
// Same ranks, same scatter targets.  The count pass only reads the kind bytes (1 B per event instead of 16); the payload of a
// tile is one contiguous word range [3*g0 + 2*c0 (+ s0), ...) known from the scanned tile counts, so the scatter stages it in
// shared memory with coalesced loads issued together with the kind loads.
// implicit != 0 (C2A_PACKED_IMPLICIT_OPERANDS): bit 7 of a gate / connection byte = "one operand is the signal declared last, no
// payload word for it"; tile_i counts those events (payload words of a tile = 3 g + 2 c - i).
__global__ void __launch_bounds__(kBlock) k_pk_count(const uint8_t* __restrict__ kinds, uint64_t n, uint32_t tiles, uint32_t implicit, uint32_t* __restrict__ tile_g,
                                                     uint32_t* __restrict__ tile_c, uint32_t* __restrict__ tile_i, uint32_t* __restrict__ es) {
  // one WARP per 1024-event tile: two coalesced 128-bit loads per lane, no block-level synchronisation
  const int lane = threadIdx.x & 31;
  const uint32_t nwarps = gridDim.x * (kBlock / 32);
  const bool aligned = !(reinterpret_cast<uintptr_t>(kinds) & 15);
  uint32_t f = 0;
  for (uint32_t tile = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5); tile < tiles; tile += nwarps) {
    const uint64_t tbase = (uint64_t)tile * kEvTile;
    uint32_t w[8];
    if (aligned && tbase + kEvTile <= n) {
      uint4 x = __ldg(reinterpret_cast<const uint4*>(kinds + tbase) + lane), y = __ldg(reinterpret_cast<const uint4*>(kinds + tbase + 512) + lane);
      w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w; w[4] = y.x; w[5] = y.y; w[6] = y.z; w[7] = y.w;
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        w[q] = 0;  // padding = SIGNAL with op 0: neither counted nor flagged
        for (int j = 0; j < 4; ++j) {
          uint64_t i = tbase + (q < 4 ? 0 : 512) + lane * 16 + (q & 3) * 4 + j;
          if (i < n) w[q] |= (uint32_t)kinds[i] << (8 * j);
        }
      }
    }
    // byte-parallel: 0/1 per byte for "gate" / "connection" / "flagged", summed over the 8 words (<= 8 per byte) and folded once;
    // the op field of all four bytes is validated with two masked compares (this kernel was ALU-bound on per-byte tests and POPC)
    uint32_t ag = 0, ac = 0, ai = 0, bad_kind = 0, bad_op = 0;
    const uint32_t opm = implicit ? 0x1F1F1F1Fu : 0x3F3F3F3Fu;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t lo = w[q] & 0x01010101u, hi = (w[q] >> 1) & 0x01010101u;  // kind bit 0 / bit 1 of each byte
      const uint32_t is_g = hi & ~lo, is_c = hi & lo;
      ag += is_g;
      ac += is_c;
      // op field (bits 2..7 of each byte, 2..6 with implicit operands): must be < 20 on a gate, 0 elsewhere (c2a_pack_events marks
      // an invalid kind that way)
      const uint32_t op = (w[q] >> 2) & opm, mg = is_g * 0xFFu;
      bad_kind |= op & ~mg;
      bad_op |= ((op & mg) + 0x01010101u * (0x80u - C2A_GATE_TYPE_COUNT)) & 0x80808080u;  // (op <= 63: no carry into the next byte)
      if (implicit) {
        const uint32_t b7 = (w[q] >> 7) & 0x01010101u;
        ai += b7 & hi;
        bad_kind |= b7 & ~hi;  // the flag on a signal declaration
      }
    }
    if (bad_kind) f |= EF_BAD_KIND;
    if (bad_op) f |= EF_BAD_OP;
    uint32_t g = (ag * 0x01010101u) >> 24, c = (ac * 0x01010101u) >> 24, im = (ai * 0x01010101u) >> 24;
    g = warp_sum(g);
    c = warp_sum(c);
    if (implicit) im = warp_sum(im);
    if (lane == 0) { tile_g[tile] = g; tile_c[tile] = c; if (implicit) tile_i[tile] = im; }
  }
  f = warp_or(f);
  if (lane == 0 && f) atomicOr(es + ES_FLAGS, f);
}

// TMA 1-D bulk copy global -> shared with completion on an mbarrier (cp.async.bulk; SASS: UBLKCP); src/dst/bytes multiples of 16
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, unsigned long long* mbar) {
  uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst), bar = (uint32_t)__cvta_generic_to_shared(mbar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}

// Persistent CTAs, two-stage shared-memory pipeline fed by TMA bulk copies: while the CTA ranks and scatters tile t, the kind
// bytes and the payload word range of its next tile are already in flight.  One elected thread computes the (16-byte aligned)
// ranges from the scanned tile counts and issues the copies; everybody waits on the stage's mbarrier.
constexpr int kPkWordsCap = 3 * kEvTile + 8;  // payload of a tile (<= 3 words per event) + alignment slack
// kMode 0: every operand spelled out.  C2A_PACKED_IMPLICIT_OPERANDS (tile_i non-null) launches BOTH other instantiations back to
// back; each looks at the scanned totals and exits at once unless the stream is its kind:
// kMode 2: every gate and connection is flagged (what the walker emits) - the third rank is the sum of the other two, a gate has 2
//          payload words and a connection 1, nothing per event has to be tested;  kMode 1: flagged and unflagged events mixed.
// kPkBlock threads per 1 024-event tile: 8 events per lane - the per-tile fixed cost (barrier wait, tile header, cross-warp prefix) is a
// quarter of the instructions of this issue-bound kernel, and it is per WARP
constexpr int kPkBlock = 128, kPkPerLane = kEvTile / kPkBlock, kPkCtasPerSm = 6;  // (7 CTAs per SM fit, at 72 registers: measured slower, 0.226 against 0.213 ms)
static_assert(kPkPerLane * kPkBlock == kEvTile && kPkBlock / 32 <= 8, "tile / block shape");
template <int kMode, bool kPoll = false>  // kPoll: the payload words are still arriving (`arrived` = the running count): wait per tile
__global__ void __launch_bounds__(kPkBlock, kPkCtasPerSm) k_pk_scatter_t(const uint8_t* __restrict__ kinds, const uint32_t* __restrict__ words, uint64_t n, uint64_t n_words,
                                                       uint32_t dense, uint32_t tiles, uint32_t S_cap, const uint32_t* __restrict__ tile_g,
                                                       const uint32_t* __restrict__ tile_c, const uint32_t* __restrict__ tile_i /* null: no implicit operands */,
                                                       uint32_t* __restrict__ sig_t, uint2* __restrict__ sig_meta,
                                                       uint4* __restrict__ egates, uint32_t* __restrict__ gate_t, uint2* __restrict__ conn,
                                                       uint32_t* __restrict__ conn_t, uint32_t* __restrict__ conn_sb, uint8_t* __restrict__ outmark,
                                                       uint32_t* __restrict__ es, const uint32_t* arrived /* null, or: payload words copied in so far */) {
  __shared__ __align__(16) uint32_t s_w[2][kPkWordsCap];
  __shared__ __align__(16) uint8_t s_k[2][kEvTile];
  __shared__ __align__(8) unsigned long long s_bar[2];
  __shared__ uint32_t s_meta[2][10];  // g0, c0, s0, first payload word - aligned start, words staged, kind bytes staged, #gates, #connections, i0, #implicit
  __shared__ uint2 s_cnt[kPkBlock / 32];  // per warp: {gates | connections << 16, implicit-operand events} of its 32 * kPkPerLane events
  __shared__ uint32_t s_list[kEvTile];  // the tile's events filed by kind (phase A -> phase B)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const bool tma_ok = !((reinterpret_cast<uintptr_t>(kinds) | reinterpret_cast<uintptr_t>(words)) & 15);
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar[b]);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // producer side (thread 0 only): counts of a tile -> ranges -> bulk copies into `stage`
  constexpr bool implicit = kMode != 0;
  if (implicit) {
    const bool all_flagged = __ldg(tile_i + tiles) == __ldg(tile_g + tiles) + __ldg(tile_c + tiles);
    if (all_flagged != (kMode == 2)) return;  // the other instantiation's stream
  }
  auto issue = [&](uint32_t tile, int stage, uint4 cnt, uint2 ci) {  // cnt = {g0, c0, g1, c1}, ci = {i0, i1}
    if (kMode == 2) ci = make_uint2(cnt.x + cnt.y, cnt.z + cnt.w);  // every gate and connection is flagged
    const uint64_t tbase = (uint64_t)tile * kEvTile, tend = min(n, tbase + kEvTile);
    const uint32_t s0 = (uint32_t)tbase - cnt.x - cnt.y, s1 = (uint32_t)tend - cnt.z - cnt.w;
    const uint64_t w0 = 3ull * cnt.x + 2ull * cnt.y - ci.x + (dense ? 0u : s0), w1 = 3ull * cnt.z + 2ull * cnt.w - ci.y + (dense ? 0u : s1);
    uint64_t a0 = w0 & ~3ull, a1 = min((w1 + 3) & ~3ull, n_words & ~3ull);
    if (a1 < a0 || !tma_ok || a1 - a0 > (uint64_t)kPkWordsCap) a1 = a0;       // (a corrupt count pair cannot overrun the stage)
    const uint32_t wbytes = (uint32_t)(a1 - a0) * 4u;
    const uint32_t kbytes = tma_ok ? (uint32_t)(tend - tbase) & ~15u : 0u;
    s_meta[stage][0] = cnt.x; s_meta[stage][1] = cnt.y; s_meta[stage][2] = s0;
    s_meta[stage][3] = (uint32_t)(w0 - a0); s_meta[stage][4] = (uint32_t)(a1 - a0); s_meta[stage][5] = kbytes;
    s_meta[stage][6] = min(cnt.z - cnt.x, (uint32_t)kEvTile); s_meta[stage][7] = min(cnt.w - cnt.y, (uint32_t)kEvTile);
    s_meta[stage][8] = ci.x; s_meta[stage][9] = min(ci.y - ci.x, (uint32_t)kEvTile);
    if (kPoll) {
      // the payload is still on its way over PCIe (chunked copy on a stream of its own, each chunk followed by a copy of the running
      // word count): wait until this tile's range has landed.  The copy engine does not depend on this kernel, so the wait ends.
      const uint32_t need = (uint32_t)min(n_words, max(a1, w1));
      uint32_t have, spins = 0;
      while (true) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(have) : "l"(arrived) : "memory");
        if (have >= need) break;
        if (++spins > (1u << 23)) __trap();  // ~8 s: the copy died - fail loudly, do not hang
        __nanosleep(1000);
      }
      asm volatile("fence.proxy.async;" ::: "memory");  // the bulk copies below read what the acquire made visible
    }
    uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar[stage]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the stage was last read through the generic proxy
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(wbytes + kbytes) : "memory");
    if (wbytes) bulk_g2s(&s_w[stage][0], words + a0, wbytes, &s_bar[stage]);
    if (kbytes) bulk_g2s(&s_k[stage][0], kinds + tbase, kbytes, &s_bar[stage]);
  };
  auto load_counts = [&](uint32_t tile) { return make_uint4(__ldg(tile_g + tile), __ldg(tile_c + tile), __ldg(tile_g + tile + 1), __ldg(tile_c + tile + 1)); };
  auto load_ci = [&](uint32_t tile) { return kMode == 1 ? make_uint2(__ldg(tile_i + tile), __ldg(tile_i + tile + 1)) : make_uint2(0u, 0u); };

  uint4 next_cnt = make_uint4(0, 0, 0, 0);
  uint2 next_ci = make_uint2(0, 0);
  uint32_t tile = blockIdx.x;
  if (threadIdx.x == 0 && tile < tiles) {
    issue(tile, 0, load_counts(tile), load_ci(tile));
    if (tile + gridDim.x < tiles) { next_cnt = load_counts(tile + gridDim.x); next_ci = load_ci(tile + gridDim.x); }
  }
  uint32_t f = 0, smax = 0;
  for (uint32_t it = 0; tile < tiles; tile += gridDim.x, ++it) {
    const int stage = it & 1;
    if (threadIdx.x == 0) {
      const uint32_t nt = tile + gridDim.x;
      if (nt < tiles) {
        issue(nt, stage ^ 1, next_cnt, next_ci);
        if (nt + gridDim.x < tiles) { next_cnt = load_counts(nt + gridDim.x); next_ci = load_ci(nt + gridDim.x); }  // consumed one iteration later
      }
    }
    {  // wait for this stage's bytes
      uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar[stage]), parity = (it >> 1) & 1u, done = 0, spins = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 24)) __trap();  // a bulk copy that never completes: fail loudly, do not hang (try_wait sleeps in hardware between polls)
      }
    }
    const uint32_t g0 = s_meta[stage][0], c0 = s_meta[stage][1], s0 = s_meta[stage][2], woff = s_meta[stage][3], wcov = s_meta[stage][4], kcov = s_meta[stage][5];
    const uint32_t ng = s_meta[stage][6], nc = s_meta[stage][7];  // gates / connections in this tile
    const uint32_t i0 = s_meta[stage][8], ni = s_meta[stage][9];  // implicit-operand events before / in this tile
    const uint64_t tbase = (uint64_t)tile * kEvTile;
    const uint32_t nev = (uint32_t)(min(n, tbase + kEvTile) - tbase);
    const uint64_t w0 = 3ull * g0 + 2ull * c0 - i0 + (dense ? 0u : s0);
    auto word = [&](uint32_t wl) -> uint32_t {  // payload word wl of this tile: staged, or (unaligned source / stream tail) straight from global
      uint32_t k = wl + woff;
      if (k < wcov) return s_w[stage][k];
      return w0 + wl < n_words ? __ldg(words + w0 + wl) : 0u;
    };
    // ---- phase A, one lane per event: rank the event inside the tile and file it under its kind.
    // s_list = [gates | connections | signals]; entry = local event index | (a second rank << 10): together with the entry's own
    // position they give all three in-tile ranks (dg + dc + ds = local index), hence the payload offset 3*dg + 2*dc (+ ds).
    // a walker stream flags EVERY gate and connection: then the third rank is the sum of the other two and need not be counted
    constexpr bool all_impl = kMode == 2;
    constexpr bool mixed = kMode == 1;
    // Interior tiles (1024 events, all kind bytes staged by the bulk copy) take the instantiation without bounds tests: this phase
    // is two thirds of the instructions of an instruction-issue-bound kernel.
    auto phase_a = [&](auto fast_tag) {
      constexpr bool kFast = decltype(fast_tag)::value;
      const uint8_t* __restrict__ skl = &s_k[stage][warp * (32 * kPkPerLane) + lane];
      uint32_t kb[kPkPerLane], gm[kPkPerLane], cm[kPkPerLane], im[kPkPerLane];
      uint32_t wg = 0, wc = 0, wi = 0;
#pragma unroll
      for (int j = 0; j < kPkPerLane; ++j) {
        if (kFast) kb[j] = skl[j * 32];
        else {
          const uint32_t k = warp * (32 * kPkPerLane) + j * 32 + lane;
          kb[j] = k < nev ? (k < kcov ? (uint32_t)s_k[stage][k] : (uint32_t)__ldg(kinds + tbase + k)) : 0x100u;  // 0x100: past the end
        }
        const bool live = kFast || kb[j] < 0x100u;
        gm[j] = __ballot_sync(0xFFFFFFFFu, live && (kb[j] & 3u) == C2A_EV_GATE);
        cm[j] = __ballot_sync(0xFFFFFFFFu, live && (kb[j] & 3u) == C2A_EV_CONNECT);
        wg += __popc(gm[j]);
        wc += __popc(cm[j]);
        if (mixed) { im[j] = __ballot_sync(0xFFFFFFFFu, live && (kb[j] & 0x82u) == 0x82u); wi += __popc(im[j]); }  // bit 7 on a gate / connection
        else im[j] = 0;
      }
      if (lane == 0) s_cnt[warp] = make_uint2(wg | (wc << 16), wi);  // (counts and their prefix sums are <= 1024: 16 bits each)
      __syncthreads();
      // in-tile ranks of the warp's first event: sum over the warps before it - lanes 0..7 hold one warp's counts each
      uint32_t dg, dc, di;
      {
        uint2 v = (lane < kPkBlock / 32 && lane < warp) ? s_cnt[lane] : make_uint2(0u, 0u);
#pragma unroll
        for (int o = 4; o; o >>= 1) { v.x += __shfl_xor_sync(0xFFFFFFFFu, v.x, o); if (mixed) v.y += __shfl_xor_sync(0xFFFFFFFFu, v.y, o); }
        v.x = __shfl_sync(0xFFFFFFFFu, v.x, 0);
        dg = v.x & 0xFFFFu;
        dc = v.x >> 16;
        di = mixed ? __shfl_sync(0xFFFFFFFFu, v.y, 0) : 0u;
      }
      // (this loop is the hot spot: one select-built shared-memory store for gates and connections, one 8-byte store for a dense
      //  signal, no per-event reductions)
      if (dense && threadIdx.x == 0 && nev > ng + nc) smax = max(smax, s0 + (nev - ng - nc));  // 1 + largest id declared in this tile
#pragma unroll
      for (int j = 0; j < kPkPerLane; ++j) {
        const uint32_t k = warp * (32 * kPkPerLane) + j * 32 + lane;
        const uint32_t my_dg = dg + __popc(gm[j] & lt), my_dc = dc + __popc(cm[j] & lt), my_di = di + __popc(im[j] & lt);
        dg += __popc(gm[j]);
        dc += __popc(cm[j]);
        di += __popc(im[j]);
        const bool is_g = (gm[j] >> lane) & 1u, is_c = (cm[j] >> lane) & 1u;
        const uint32_t ds = k - my_dg - my_dc;
        if (is_g | is_c) {  // (both masks exclude lanes past the end of the stream)
          // entry: local event index | the rank under the other kind << 10 | the rank among implicit-operand events << 20
          const uint32_t slot = is_g ? my_dg : ng + my_dc;
          s_list[kFast ? slot : min(slot, (uint32_t)kEvTile - 1)] = k | ((is_g ? my_dc : my_dg) << 10) | (my_di << 20);
        } else if (kFast || k < nev) {
          const uint32_t cbit = (kb[j] & 3u) == C2A_EV_SIGNAL_CONST ? 0x80000000u : 0u;
          // dense ids: the id IS the declaration rank (< n <= S_cap) - the record is complete right here, lanes holding signals write
          // consecutive slots (no filing, no phase-B pass for the most frequent kind)
          if (dense) reinterpret_cast<uint32_t*>(sig_meta)[s0 + ds] = (c0 + my_dc) | cbit;  // dense record: 4 bytes (the rank is the index)
          else s_list[min(ng + nc + ds, (uint32_t)kEvTile - 1)] = k | (my_dg << 10) | cbit;
        }
      }
    };
    // ---- dense ids, interior tile: ONE pass, one lane per event, every record stored from the lane that ranked it.  The lanes of a
    // warp split three ways (gate / connection / signal), but nothing is filed in shared memory and re-derived in a second pass -
    // about half the instructions of phase A + phase B in a kernel that is bound by instruction issue.  Lanes that hold the same
    // kind write consecutive records, so the stores still fill whole sectors.
    auto phase_direct = [&]() {
      const uint8_t* __restrict__ skl = &s_k[stage][warp * (32 * kPkPerLane) + lane];
      const uint32_t* __restrict__ sw = &s_w[stage][woff];
      uint32_t kb[kPkPerLane], gm[kPkPerLane], cm[kPkPerLane], im[kPkPerLane];
      uint32_t wg = 0, wc = 0, wi = 0;
#pragma unroll
      for (int j = 0; j < kPkPerLane; ++j) {
        kb[j] = skl[j * 32];
        gm[j] = __ballot_sync(0xFFFFFFFFu, (kb[j] & 3u) == C2A_EV_GATE);
        cm[j] = __ballot_sync(0xFFFFFFFFu, (kb[j] & 3u) == C2A_EV_CONNECT);
        wg += __popc(gm[j]);
        wc += __popc(cm[j]);
        if (mixed) { im[j] = __ballot_sync(0xFFFFFFFFu, (kb[j] & 0x82u) == 0x82u); wi += __popc(im[j]); }
        else im[j] = 0;
      }
      if (lane == 0) s_cnt[warp] = make_uint2(wg | (wc << 16), wi);
      __syncthreads();
      uint32_t dg, dc, di;
      {
        uint2 v = (lane < kPkBlock / 32 && lane < warp) ? s_cnt[lane] : make_uint2(0u, 0u);
#pragma unroll
        for (int o = 4; o; o >>= 1) { v.x += __shfl_xor_sync(0xFFFFFFFFu, v.x, o); if (mixed) v.y += __shfl_xor_sync(0xFFFFFFFFu, v.y, o); }
        v.x = __shfl_sync(0xFFFFFFFFu, v.x, 0);
        dg = v.x & 0xFFFFu;
        dc = v.x >> 16;
        di = mixed ? __shfl_sync(0xFFFFFFFFu, v.y, 0) : 0u;
      }
      if (threadIdx.x == 0 && nev > ng + nc) smax = max(smax, s0 + (nev - ng - nc));  // 1 + largest id declared in this tile
#pragma unroll
      for (int j = 0; j < kPkPerLane; ++j) {
        const uint32_t k = warp * (32 * kPkPerLane) + j * 32 + lane;
        const uint32_t my_dg = dg + __popc(gm[j] & lt), my_dc = dc + __popc(cm[j] & lt);
        const uint32_t my_di = all_impl ? my_dg + my_dc : (mixed ? di + __popc(im[j] & lt) : 0u);
        dg += __popc(gm[j]);
        dc += __popc(cm[j]);
        if (mixed) di += __popc(im[j]);
        const uint32_t before = s0 + (k - my_dg - my_dc);  // signals declared before this event (dense ids: every id below it exists)
        const uint32_t wl = 3u * my_dg + 2u * my_dc - my_di;
        const bool flagged = all_impl || (implicit && (kb[j] & 0x80u));
        if ((gm[j] >> lane) & 1u) {
          uint4 gt = make_uint4(implicit ? (kb[j] >> 2) & 31u : kb[j] >> 2, sw[wl], sw[wl + 1], flagged ? before - 1u : sw[wl + 2]);
          if (gt.y < before && gt.z < before && gt.w < before) outmark[gt.w] = 1;  // compiler.rs:201 marks the out node is_out
          else { f |= EF_UNKNOWN_REF; gt.y = gt.z = gt.w = 0; }
          egates[g0 + my_dg] = gt;
        } else if ((cm[j] >> lane) & 1u) {
          uint2 ab = flagged ? make_uint2(before - 1u, sw[wl]) : make_uint2(sw[wl], sw[wl + 1]);
          if (!(ab.x < before && ab.y < before)) { f |= EF_UNKNOWN_REF; ab = make_uint2(0, 0); }
          conn[c0 + my_dc] = ab;
          conn_sb[c0 + my_dc] = before;
        } else {
          reinterpret_cast<uint32_t*>(sig_meta)[before] = (c0 + my_dc) | ((kb[j] & 3u) == C2A_EV_SIGNAL_CONST ? 0x80000000u : 0u);
        }
      }
    };
    const uint32_t wn_all = 3u * ng + 2u * nc - ni;  // payload words of a dense tile
    if (dense && nev == (uint32_t)kEvTile && kcov == (uint32_t)kEvTile && ng + nc <= (uint32_t)kEvTile && woff + wn_all <= wcov) {
      phase_direct();
      __syncthreads();  // the stage may be refilled from the next iteration on
      continue;
    }
    if (nev == (uint32_t)kEvTile && kcov == (uint32_t)kEvTile && ng + nc <= (uint32_t)kEvTile) phase_a(std::true_type{});
    else phase_a(std::false_type{});
    __syncthreads();
    // ---- phase B, one lane per OUTPUT record, kind by kind: no divergence, fully coalesced stores.
    // Interior tiles (everything staged by the bulk copies) take the check-free instantiation.
    auto phase_b = [&](auto fast_tag) {
      constexpr bool kFast = decltype(fast_tag)::value;
      const uint32_t* __restrict__ sw = &s_w[stage][woff];
      const uint8_t* __restrict__ sk = &s_k[stage][0];
      auto W = [&](uint32_t wl) -> uint32_t { return kFast ? sw[wl] : word(wl); };
      auto K = [&](uint32_t k) -> uint32_t { return kFast ? (uint32_t)sk[k] : (k < kcov ? (uint32_t)sk[k] : (uint32_t)__ldg(kinds + tbase + k)); };
      // Dense ids (id = declaration rank): "declared before use" (compiler.rs:183, :201 resolve an undeclared signal to node 0 /
      // panic) is the local test id < #signals declared before this event, so the E2 kernels, the event-time arrays and the
      // declaration table are not needed at all; out-of-range references are flagged and neutralised right here.
      for (uint32_t r = threadIdx.x; r < ng; r += kPkBlock) {  // gate r of the tile
        uint32_t e = s_list[r], k = e & 1023u, my_dc = (e >> 10) & 1023u, my_di = all_impl ? r + my_dc : e >> 20;
        uint32_t ds = k - r - my_dc;
        uint32_t wl = 3u * r + 2u * my_dc - my_di + (dense ? 0u : ds);
        const uint32_t kbyte = K(k);
        const bool im_out = all_impl || (implicit && (kbyte & 0x80u));  // out = the signal declared last (dense ids: its id is its rank)
        uint4 gt = make_uint4(implicit ? (kbyte >> 2) & 31u : kbyte >> 2, W(wl), W(wl + 1), im_out ? s0 + ds - 1u : W(wl + 2));
        if (dense) {
          const uint32_t before = s0 + ds;
          if (gt.y < before && gt.z < before && gt.w < before) outmark[gt.w] = 1;  // compiler.rs:201 marks the out node is_out
          else { f |= EF_UNKNOWN_REF; gt.y = gt.z = gt.w = 0; }
        } else gate_t[g0 + r] = (uint32_t)tbase + k;
        egates[g0 + r] = gt;
      }
      for (uint32_t r = threadIdx.x; r < nc; r += kPkBlock) {  // connection r of the tile
        uint32_t e = s_list[ng + r], k = e & 1023u, my_dg = (e >> 10) & 1023u, my_di = all_impl ? my_dg + r : e >> 20;
        uint32_t ds = k - my_dg - r;
        uint32_t wl = 3u * my_dg + 2u * r - my_di + (dense ? 0u : ds);
        const bool im_a = all_impl || (implicit && (K(k) & 0x80u));  // a = the signal declared last
        uint2 ab = im_a ? make_uint2(s0 + ds - 1u, W(wl)) : make_uint2(W(wl), W(wl + 1));
        if (dense) {
          if (!(ab.x < s0 + ds && ab.y < s0 + ds)) { f |= EF_UNKNOWN_REF; ab = make_uint2(0, 0); }
        } else conn_t[c0 + r] = (uint32_t)tbase + k;
        conn[c0 + r] = ab;
        conn_sb[c0 + r] = s0 + ds;  // signals declared before the connection
      }
      const uint32_t ns = dense ? 0u : nev - min(nev, ng + nc);  // (dense: already stored in phase A)
      for (uint32_t r = threadIdx.x; r < ns; r += kPkBlock) {  // signal r of the tile
        uint32_t e = s_list[ng + nc + r], k = e & 1023u, my_dg = (e >> 10) & 1023u, my_dc = k - my_dg - r;
        uint32_t sid = dense ? s0 + r : W(3u * my_dg + 2u * my_dc + r);
        if (sid == 0xFFFFFFFFu) f |= EF_SPARSE;
        else {
          smax = max(smax, sid + 1);
          if (sid >= S_cap) f |= EF_CAP;
          else {
            // plain stores: a duplicate declaration (compiler.rs:146-148) overwrites, and is caught later because the number of
            // declared ids then falls short of the number of signal events (k_ev_finalize counts them)
            if (!dense) sig_t[sid] = (uint32_t)tbase + k;
            sig_meta[sid] = make_uint2((s0 + r) | (e & 0x80000000u), c0 + my_dc);
          }
        }
      }
    };
    const uint32_t wn = 3u * ng + 2u * nc - ni + (dense ? 0u : nev - min(nev, ng + nc));  // payload words of the tile
    if (kcov == nev && woff + wn <= wcov && ng + nc <= nev) phase_b(std::true_type{});
    else phase_b(std::false_type{});
    __syncthreads();  // the stage (and s_g / s_c) may be refilled from the next iteration on
  }
  smax = warp_max(smax);
  f = warp_or(f);
  if (lane == 0) {
    if (smax) atomicMax(es + ES_SBOUND, smax);
    if (f) atomicOr(es + ES_FLAGS, f);
  }
}

// ---- E2 ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_ev_check_gates(uint4* __restrict__ egates, const uint32_t* __restrict__ gate_t, uint32_t G, uint32_t S,
                                                           const uint32_t* __restrict__ sig_t, uint8_t* __restrict__ outmark, uint32_t* __restrict__ es) {
  bool bad = false;
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < G; g += gridDim.x * kBlock) {
    uint4 e = egates[g];
    uint32_t t = gate_t[g];
    if (!(e.y < S && e.z < S && e.w < S)) {
      // out-of-range reference: flagged, and neutralised so that every later kernel stays memory-safe; the flags are only
      // looked at once, at the end (the stream is then replayed by the host emitter)
      bad = true;
      egates[g] = make_uint4(e.x, 0, 0, 0);
      continue;
    }
    // a signal that is not declared BEFORE the gate resolves to node 0 / panics in the reference (compiler.rs:183, :201)
    bool ok = __ldg(sig_t + e.y) < t && __ldg(sig_t + e.z) < t && __ldg(sig_t + e.w) < t;
    if (!ok) bad = true;
    else outmark[e.w] = 1;  // compiler.rs:201 marks the out node is_out
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(es + ES_FLAGS, (uint32_t)EF_UNKNOWN_REF);
}
__global__ void __launch_bounds__(kBlock) k_ev_check_conns(uint2* __restrict__ conn, const uint32_t* __restrict__ conn_t, uint32_t C, uint32_t S,
                                                           const uint32_t* __restrict__ sig_t, uint32_t* __restrict__ es) {
  bool bad = false;
  for (uint32_t c = blockIdx.x * kBlock + threadIdx.x; c < C; c += gridDim.x * kBlock) {
    uint2 ab = conn[c];
    uint32_t t = conn_t[c];
    if (!(ab.x < S && ab.y < S)) { bad = true; conn[c] = make_uint2(0, 0); continue; }
    if (!(__ldg(sig_t + ab.x) < t && __ldg(sig_t + ab.y) < t)) bad = true;
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(es + ES_FLAGS, (uint32_t)EF_UNKNOWN_REF);
}

// ---- M: Boruvka over the connection graph, weight = connection index (= event order) -------------------------
__device__ __forceinline__ uint32_t uf_find(uint32_t* __restrict__ parent, uint32_t x) {
  // Roots do not change while a kernel that calls this runs (hooking happens in k_msf_hook only); the compression
  // stores race benignly: every value written is the current root.
  uint32_t r = x, p;
  while ((p = parent[r]) != r) r = p;
  while (x != r) { p = parent[x]; if (p != r) parent[x] = r; x = p; }
  return r;
}


__device__ __forceinline__ void red_min_u32(uint32_t* p, uint32_t v) {
  if (*reinterpret_cast<volatile uint32_t*>(p) > v) atomicMin(p, v);  // values only decrease within a round: a stale read costs one RED
}

// First round: the live list is every connection and the classes are the signals themselves, so a candidate record would
// only repeat conn[e]: nothing is materialised, k_msf_hook_first reads conn[] again.
__global__ void __launch_bounds__(kBlock) k_msf_pick_first(const uint2* __restrict__ conn, uint32_t n, uint32_t* __restrict__ best, uint32_t tag,
                                                           uint32_t* __restrict__ n_cand) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_cand = n;
  for (uint32_t e = blockIdx.x * kBlock + threadIdx.x; e < n; e += gridDim.x * kBlock) {
    uint2 ab = conn[e];
    if (ab.x != ab.y) {  // parent[] is still the identity: the classes are the signals themselves
      uint32_t val = tag | e;
      red_min_u32(best + ab.x, val);
      red_min_u32(best + ab.y, val);
    }
  }
}

// Later rounds: cur[0..*n_cur) is the live list.  cand[] receives {edge, class of a, class of b, 0}, compacted.
__global__ void __launch_bounds__(kBlock) k_msf_pick(const uint2* __restrict__ conn, const uint32_t* __restrict__ cur, const uint32_t* __restrict__ n_cur,
                                                     uint32_t* __restrict__ parent, uint32_t* __restrict__ best, uint32_t tag,
                                                     uint4* __restrict__ cand, uint32_t* __restrict__ n_cand) {
  const uint32_t n = *n_cur;
  const int lane = threadIdx.x & 31;
  for (uint32_t i0 = blockIdx.x * kBlock + (threadIdx.x & ~31u); i0 < n; i0 += gridDim.x * kBlock) {
    uint32_t i = i0 + lane;
    bool keep = false;
    uint4 out = make_uint4(0, 0, 0, 0);
    if (i < n) {
      uint32_t e = cur[i];
      uint2 ab = conn[e];
      uint32_t cu = uf_find(parent, ab.x), cv = uf_find(parent, ab.y);
      if (cu != cv) {  // still joins two classes: candidate; otherwise it is (or became) internal and is dropped
        uint32_t val = tag | e;
        red_min_u32(best + cu, val);
        red_min_u32(best + cv, val);
        keep = true;
        out = make_uint4(e, cu, cv, 0);
      }
    }
    warp_append(keep, out, cand, n_cand);
  }
}

// A class hooks along its minimum edge (an MSF edge => an effective connection).  Mutual minimum: the larger root
// goes under the smaller.  Edges not chosen by either side stay on the live list.
template <bool kFirst>  // kFirst: candidates are conn[0..n) themselves (round 1); otherwise cand[0..*n_cand)
__global__ void __launch_bounds__(kBlock) k_msf_hook_t(const uint4* __restrict__ cand, const uint2* __restrict__ conn, const uint32_t* __restrict__ n_cand,
                                                       uint32_t* __restrict__ parent, const uint32_t* __restrict__ best, uint32_t tag, uint32_t* __restrict__ eff,
                                                       uint32_t* __restrict__ cur, uint32_t* __restrict__ n_cur, uint32_t* __restrict__ n_rounds) {
  const uint32_t n = *n_cand;
  const int lane = threadIdx.x & 31;
  if (n && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(n_rounds, 1u);  // a round that examined candidates
  for (uint32_t i0 = blockIdx.x * kBlock + (threadIdx.x & ~31u); i0 < n; i0 += gridDim.x * kBlock) {
    uint32_t i = i0 + lane;
    bool keep = false;
    uint32_t e = 0;
    uint4 c = make_uint4(kNone, 0, 0, 0);
    if (i < n) {
      if (kFirst) {
        uint2 ab = conn[i];
        if (ab.x != ab.y) c = make_uint4(i, ab.x, ab.y, 0);
      } else c = cand[i];
    }
    if (c.x != kNone) {
      e = c.x;
      uint32_t val = tag | e;
      bool bu = best[c.y] == val, bv = best[c.z] == val;
      if (bu && bv) parent[max(c.y, c.z)] = min(c.y, c.z);
      else if (bu) parent[c.y] = c.z;
      else if (bv) parent[c.z] = c.y;
      else keep = true;
      if (!kFirst && !keep) atomicOr(eff + (e >> 5), 1u << (e & 31));  // effective connection: one bit (the bitmap is L2-resident)
    }
    if (kFirst) {  // round 1: the warp's 32 connections are exactly one bitmap word - one store, no atomics
      uint32_t m = __ballot_sync(0xFFFFFFFFu, c.x != kNone && !keep);
      if (lane == 0 && m) atomicOr(eff + (i0 >> 5), m);
    }
    warp_append(keep, e, cur, n_cur);
  }
}

// ---- N: node ids ----------------------------------------------------------------------------------------------
// nidf[root] = largest id among the class's effective connections (0 = none) in the low 30 bits (ids < 2^29: n <= 2^29 events);
// bits 31 / 30 are the merge screens "the class already holds a constant / a gate output" (set in k_ev_finalize, after every
// atomicMax of k_ev_nid_edges has landed).  One 4-byte word per signal: the gathers of both kernels touch half the sectors an
// {id, counters} pair did - these kernels are bound by L1 wavefronts (one per distinct sector per warp instruction), not by DRAM.
// The ids grow with the
// event index, and every member of a class of >= 2 signals is joined by an effective connection AFTER its declaration, so the
// class ends with the id of its last effective connection (compiler.rs:257); a class without one is a single signal and keeps
// the id of its declaration (compiler.rs:157).
// Pointer chases are latency chains (coalesced loads -> parent[a] -> parent[root] -> atomic); every thread keeps kChaseIlp of
// them in flight and probes the second hop speculatively (after one Boruvka round every tree has depth 1).
// eff[] is a bitmap over the connections, effp[w] the number of effective connections before word w (k_scan_u32_t<popc>):
// #effective connections before connection c
__device__ __forceinline__ uint32_t eff_rank(const uint32_t* __restrict__ eff, const uint32_t* __restrict__ effp, uint32_t c) {
  return __ldg(effp + (c >> 5)) + __popc(__ldg(eff + (c >> 5)) & ((1u << (c & 31)) - 1u));
}
constexpr int kChaseIlp = 4;
constexpr uint32_t kHasConst = 0x80000000u, kHasOut = 0x40000000u, kNidMask = 0x3FFFFFFFu;
__global__ void __launch_bounds__(kBlock) k_ev_nid_edges(uint32_t C, const uint2* __restrict__ conn, const uint32_t* __restrict__ conn_sb,
                                                         const uint32_t* __restrict__ eff, const uint32_t* __restrict__ effp,
                                                         uint32_t* __restrict__ parent, uint32_t* __restrict__ nidf) {
  const uint32_t stride = gridDim.x * kBlock;
  const int lane = threadIdx.x & 31;
  for (uint32_t w0 = blockIdx.x * kBlock + (threadIdx.x & ~31u); w0 < C; w0 += stride * kChaseIlp) {  // warp-uniform trip count (match below)
    const uint32_t c0 = w0 + lane;
    uint32_t x[kChaseIlp], sb[kChaseIlp], a[kChaseIlp], r0[kChaseIlp], r1[kChaseIlp];
    bool is_eff[kChaseIlp];
#pragma unroll
    for (int i = 0; i < kChaseIlp; ++i) {
      uint32_t c = min(c0 + i * stride, C - 1);
      uint32_t w = __ldg(eff + (c >> 5));
      is_eff[i] = (w >> (c & 31)) & 1u;
      x[i] = __ldg(effp + (c >> 5)) + __popc(w & ((1u << (c & 31)) - 1u));
      sb[i] = conn_sb[c];
      a[i] = conn[c].x;
    }
#pragma unroll
    for (int i = 0; i < kChaseIlp; ++i) r0[i] = parent[a[i]];
#pragma unroll
    for (int i = 0; i < kChaseIlp; ++i) r1[i] = r0[i] == a[i] ? r0[i] : parent[r0[i]];  // a root needs no second probe (fewer sectors per warp)
    // One class may hold millions of effective connections (a key wired into every round of every chain: 1.7 M on the MiMC program
    // compiled from source - 1.8 ms of same-address atomics).  Ids grow with the connection index, so of the lanes of a warp that
    // share a root only the HIGHEST lane needs the atomic, and a word that already holds a larger id needs none.
#pragma unroll
    for (int i = 0; i < kChaseIlp; ++i) {
      const bool act = c0 + i * stride < C && is_eff[i];      // not effective: no id consumed (compiler.rs:235-237)
      const uint32_t id = sb[i] + x[i] + 1u;                  // compiler.rs:257 with node_count = signals + effective merges so far
      const uint32_t r = !act ? kNone : (r1[i] == r0[i] ? r0[i] : uf_find(parent, a[i]));
      const uint32_t peers = __match_any_sync(0xFFFFFFFFu, r);
      if (act && 31 - __clz(peers) == lane && (__ldcg(nidf + r) & kNidMask) < id) atomicMax(nidf + r, id);
    }
  }
}
// node_of_signal + the merge-error screens (compiler.rs:239-245)
constexpr int kFinIlp = 2;
__global__ void __launch_bounds__(kBlock) k_ev_finalize(uint32_t S, const uint32_t* __restrict__ sig_t, const uint2* __restrict__ sig_meta,
                                                        const uint8_t* __restrict__ outmark, const uint32_t* __restrict__ eff, const uint32_t* __restrict__ effp,
                                                        uint32_t* __restrict__ parent, uint32_t* __restrict__ nidf, uint32_t* __restrict__ nos,
                                                        uint32_t* __restrict__ es) {
  uint32_t f = 0, declared = 0;
  const uint32_t stride = gridDim.x * kBlock;
  for (uint32_t s0 = blockIdx.x * kBlock + threadIdx.x; s0 < S; s0 += stride * kFinIlp) {
    uint32_t t[kFinIlp], r0[kFinIlp], r1[kFinIlp], nid[kFinIlp], om[kFinIlp];
    uint2 m[kFinIlp];
#pragma unroll
    for (int i = 0; i < kFinIlp; ++i) {
      uint32_t s = min(s0 + i * stride, S - 1);
      t[i] = sig_t ? sig_t[s] : 0u;  // dense ids: no declaration table, every id below S is declared
      if (!sig_t) {  // dense record: #connections before the declaration | is_const << 31; the declaration rank is the id
        const uint32_t v = reinterpret_cast<const uint32_t*>(sig_meta)[s];
        m[i] = make_uint2(s | (v & 0x80000000u), v & 0x7FFFFFFFu);
      } else m[i] = t[i] != kNone ? sig_meta[s] : make_uint2(0, 0);  // (an undeclared id has no record: do not read uninitialised memory)
      om[i] = outmark[s];
      r0[i] = parent[s];
    }
#pragma unroll
    for (int i = 0; i < kFinIlp; ++i) {
      r1[i] = r0[i] == min(s0 + i * stride, S - 1) ? r0[i] : parent[r0[i]];  // a root needs no second probe
      nid[i] = __ldcg(nidf + r0[i]) & kNidMask;
    }
#pragma unroll
    for (int i = 0; i < kFinIlp; ++i) {
      uint32_t s = s0 + i * stride;
      if (s >= S) continue;
      uint32_t node = 0;
      if (t[i] != kNone) {
        ++declared;
        uint32_t r = r0[i];
        node = nid[i];
        if (r1[i] != r0[i]) { r = uf_find(parent, s); node = __ldcg(nidf + r) & kNidMask; }
        if (node == 0) {
          node = (m[i].x & 0x7FFFFFFFu) + 1u + eff_rank(eff, effp, m[i].y);  // a class of one: compiler.rs:157
        } else {  // merged class: at most one constant and one gate output may meet in it (compiler.rs:239-245)
          if (m[i].x & 0x80000000u) { if (atomicOr(nidf + r, kHasConst) & kHasConst) f |= EF_CONST_CONST; }
          if (om[i]) { if (atomicOr(nidf + r, kHasOut) & kHasOut) f |= EF_OUT_OUT; }
        }
      }
      nos[s] = node;
    }
  }
  f = warp_or(f);
  declared = warp_sum(declared);
  if ((threadIdx.x & 31) == 0) {
    if (f) atomicOr(es + ES_FLAGS, f);
    if (declared) atomicAdd(es + ES_NDECL, declared);
  }
}
// Also K1 of the build that follows (compiler.rs:401-406): prod1[out node] = 1 + last gate writing it, while the node ids are in
// registers - the build then neither zeroes prod1 nor reads the gate array a second time for k_producer.
__global__ void __launch_bounds__(kBlock) k_ev_gates(const uint4* __restrict__ egates, uint32_t G, uint32_t S, const uint32_t* __restrict__ nos,
                                                     uint4* __restrict__ gates, uint32_t* __restrict__ prod1, uint32_t prod_bound) {
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < G; g += gridDim.x * kBlock) {
    uint4 e = egates[g];
    if (S == 0) {  // only on a stream that is about to be declined
      stg_stream(gates + g, make_uint4(e.x, 0, 0, 0));
      continue;
    }
    const uint32_t o = __ldg(nos + e.w);
    stg_stream(gates + g, make_uint4(e.x, __ldg(nos + e.y), __ldg(nos + e.z), o));
    if (o < prod_bound) atomicMax(prod1 + o, g + 1);  // RED.MAX, no return value
  }
}
// I/O signal ids -> node ids (compiler.rs:327-361 walks nodes; here the caller lists signals)
__global__ void __launch_bounds__(kBlock) k_ev_map_io(const uint32_t* __restrict__ sigs, uint32_t n, uint32_t S, const uint32_t* __restrict__ nos,
                                                      uint32_t* __restrict__ nodes, uint32_t* __restrict__ es) {
  bool bad = false;
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    uint32_t s = sigs[i], nd = s < S ? nos[s] : 0u;
    if (nd == 0) bad = true;
    nodes[i] = nd;
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) es[ES_IOBAD] = 1u;
}

// selected signals -> wire ids (name-map lookups of compiler.rs:323-383, 466-493)
__global__ void __launch_bounds__(kBlock) k_sig_wires(const uint32_t* __restrict__ sigs, uint64_t n, uint32_t S, const uint32_t* __restrict__ nos,
                                                      const uint32_t* __restrict__ wire, uint32_t node_bound, uint32_t* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kBlock) {
    uint32_t s = sigs[i];
    uint32_t nd = nos ? (s < S ? nos[s] : 0u) : s;  // nos == nullptr: the list already holds node ids
    out[i] = wire ? (nd && nd < node_bound ? wire[nd] : kNone) : nd;  // wire == nullptr: the node ids themselves
  }
}

// ---------------------------------------------------------------------------------------------------------------
static inline size_t emit_scratch_bytes(uint64_t G, uint64_t C, uint64_t S) {  // slab part (exact sizes; the scatter targets live in the staging buffer)
  size_t b = 0;
  b += 3 * align256(4 * S);                                                 // parent, best, nidf
  b += 2 * align256(4 * (C / 32 + 4)) + align256(4 * C) + align256(16 * C); // eff bitmap, its rank prefix, cur, cand
  b += align256(8 * (size_t)(scan_tiles(C + 1, kScanItems) + 1)) + 256;     // tile_state + ticket
  (void)G;
  return b;
}

static void emit_flags_text(uint32_t f, char* buf, size_t cap) {
  snprintf(buf, cap, "%s%s%s%s%s%s%s", (f & EF_BAD_KIND) ? "bad-kind " : "", (f & EF_BAD_OP) ? "bad-op " : "", (f & EF_DUPLICATE) ? "duplicate-signal " : "",
           (f & EF_UNKNOWN_REF) ? "reference-before-declaration " : "", (f & EF_CONST_CONST) ? "const+const-merge " : "",
           (f & EF_OUT_OUT) ? "out+out-merge? " : "", (f & EF_SPARSE) ? "sparse-signal-ids " : "");
}

}  // namespace c2a

using namespace c2a;

extern "C" {
// host emitter entry points (c2a_host.cpp) used for the exact replay of streams the device path declines
c2a_compiler* c2a_compiler_new(void);
void c2a_compiler_free(c2a_compiler*);
int c2a_emit_events(c2a_compiler*, const c2a_event* ev, uint64_t n, uint64_t* err_event);
uint64_t c2a_num_gates(const c2a_compiler*);
uint32_t c2a_node_count(const c2a_compiler*);
uint64_t c2a_num_signals(const c2a_compiler*);
int c2a_get_gates(c2a_compiler*, c2a_gate* out);
int c2a_signal_nodes(c2a_compiler*, const uint32_t* signal_ids, uint64_t n, uint32_t* node_ids);
const char* c2a_compiler_last_error(const c2a_compiler*);
}

namespace c2a {

static void emit_drop_host(c2a_handle* h) {
  if (h->host_comp) { c2a_compiler_free(h->host_comp); h->host_comp = nullptr; }
}

// Exact replay on the host emitter; uploads the node-id gate vector (and node_of_signal when ids are dense).
static int emit_host_path(c2a_handle* h, const c2a_event* ev, uint64_t n, uint32_t sbound_hint, bool sparse, c2a_emit_info* info, uint64_t* err_event) {
  emit_drop_host(h);
  c2a_compiler* c = c2a_compiler_new();
  uint64_t bad = 0;
  int st = c2a_emit_events(c, ev, n, &bad);
  if (st != C2A_OK) {
    if (err_event) *err_event = bad;
    fail(h, st, "event %llu: %s", (unsigned long long)bad, st == C2A_ERR_INVALID_ARGUMENT || st == C2A_ERR_REFERENCE_PANIC ? c2a_compiler_last_error(c) : c2a_status_string(st));
    c2a_compiler_free(c);
    return st;
  }
  const uint64_t G = c2a_num_gates(c);
  const uint32_t S = sparse ? 0u : sbound_hint;
  std::vector<c2a_gate> gates(G);
  c2a_get_gates(c, gates.data());
  std::vector<uint32_t> nos(S), ids(S);
  for (uint32_t i = 0; i < S; ++i) ids[i] = i;
  if (S) c2a_signal_nodes(c, ids.data(), S, nos.data());
  slab_reset(h);
  size_t resident = align256(16 * G) + align256(4 * (size_t)S);
  if (!slab_reserve(h, resident)) { c2a_compiler_free(c); return C2A_ERR_NO_MEMORY; }
  uint4* d_gates = (uint4*)slab_alloc(h, 16 * G);
  uint32_t* d_nos = (uint32_t*)slab_alloc(h, 4 * (size_t)S);
  if (G) cudaMemcpyAsync(d_gates, gates.data(), 16 * G, cudaMemcpyHostToDevice, h->stream);
  if (S) cudaMemcpyAsync(d_nos, nos.data(), 4 * (size_t)S, cudaMemcpyHostToDevice, h->stream);
  if (!cuda_ok(h, cudaStreamSynchronize(h->stream), "host-path upload")) { c2a_compiler_free(c); return C2A_ERR_CUDA; }
  h->slab_keep = h->slab_used;
  h->emitted.valid = true;
  h->emitted.nos_valid = !sparse;
  h->emitted.gates_off = (char*)d_gates - h->slab;
  h->emitted.nos_off = (char*)d_nos - h->slab;
  h->emitted.G = G;
  h->emitted.node_count = c2a_node_count(c);
  h->emitted.signal_bound = S;
  h->emitted.wire = nullptr;
  h->host_comp = c;  // kept: I/O signal lists are mapped through it when node_of_signal is not resident
  if (info) {
    info->n_gates = G;
    info->n_signals = c2a_num_signals(c);
    info->n_effective = h->emitted.node_count - info->n_signals;
    info->node_count = h->emitted.node_count;
    info->signal_bound = S;
    info->path = C2A_EMIT_PATH_HOST;
  }
  return C2A_OK;
}

// ---- compressed streams (include/c2a.h: c2a_compressed_events) ---------------------------------------------------------------
// The host splits literal ranges and replay records into chunks of at most kCxChunk elements; one CTA per chunk copies it
// (payload words: + delta).  Chunks of one generation are independent; generations run as consecutive launches.
constexpr uint32_t kCxChunk = 16384;
struct CxChunk {
  unsigned long long dst, src;  // element offsets (bytes for the kind array, words for the payload)
  uint32_t len, delta;
};
__global__ void __launch_bounds__(kBlock) k_cx_copy_u8(const CxChunk* __restrict__ chunks, uint32_t n_chunks, const uint8_t* src_base, uint8_t* dst_base) {
  for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const CxChunk ch = chunks[c];
    const uint8_t* s = src_base + ch.src;
    uint8_t* d = dst_base + ch.dst;
    if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 3) == 0) {  // word-wise body, byte tail
      const uint32_t nw = ch.len >> 2;
      for (uint32_t i = threadIdx.x; i < nw; i += kBlock) reinterpret_cast<uint32_t*>(d)[i] = reinterpret_cast<const uint32_t*>(s)[i];
      for (uint32_t i = (nw << 2) + threadIdx.x; i < ch.len; i += kBlock) d[i] = s[i];
    } else
      for (uint32_t i = threadIdx.x; i < ch.len; i += kBlock) d[i] = s[i];
  }
}
__global__ void __launch_bounds__(kBlock) k_cx_copy_u32(const CxChunk* __restrict__ chunks, uint32_t n_chunks, const uint32_t* src_base, uint32_t* dst_base) {
  for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const CxChunk ch = chunks[c];
    const uint32_t* s = src_base + ch.src;
    uint32_t* d = dst_base + ch.dst;
    for (uint32_t i = threadIdx.x; i < ch.len; i += kBlock) d[i] = s[i] + ch.delta;
  }
}

}  // namespace c2a

extern "C" {

// ev_host: events in host memory (copied in) or, when ev_dev != null, already resident on the handle's device.
// Source of the stream: AoS events (host or device pointer) or a packed stream (kinds / words on the host or on the device).
struct EmitSrc {
  const c2a_event* ev_host = nullptr;
  const c2a_event* ev_dev = nullptr;
  const c2a_packed_events* pk = nullptr;
  bool pk_on_device = false;
  bool keep_phases = false;  // the caller already opened this call's phase list (the expansion of a compressed stream)
};

// c2a_compile_packed* on a large stream: the emit does not wait for its own final status - the build is enqueued right behind
// it (with an upper bound as node bound) and ONE synchronisation at the end of the build returns both.  What the emit could
// not check is validated then (compile_packed_impl); a stream that fails the check is replayed the ordinary way.
struct EmitDefer {
  bool pending = false;
  uint32_t wire_cap = 0;  // capacity of the caller's device wire map (0: none) - caps the provisional node bound
  uint64_t n_sig = 0, C = 0;
  uint32_t* es = nullptr;  // the emit's device status block (stays valid through the build: ES_IOBAD is the build's to use)
};

static int emit_events_impl(c2a_handle* h, const EmitSrc& src, uint64_t n, c2a_emit_info* info, uint64_t* err_event, EmitDefer* defer = nullptr) {
  const c2a_event* ev_host = src.ev_host;
  const c2a_event* ev_dev = src.ev_dev;
  const c2a_packed_events* pk = src.pk;
  const bool pk_dense = pk && (pk->flags & C2A_PACKED_DENSE_IDS);
  const bool pk_impl = pk_dense && (pk->flags & C2A_PACKED_IMPLICIT_OPERANDS);  // bit 7 of a gate / connection byte: one operand is implicit
  int st = check_sizes(h, n, 1);
  if (st) return st;
  if (n && !ev_host && !ev_dev && !pk) return fail(h, C2A_ERR_INVALID_ARGUMENT, "null event array");
  if (pk && n && (!pk->kinds || (pk->n_words && !pk->words))) return fail(h, C2A_ERR_INVALID_ARGUMENT, "null packed arrays");
  if (pk && pk->n_words > 3 * n) return fail(h, C2A_ERR_INVALID_ARGUMENT, "n_words does not match the kinds");
  const c2a_event* ev = ev_host;
  if (info) memset(info, 0, sizeof *info);
  if (info) info->n_events = n;
  if (!src.keep_phases) phases_clear(h);
  cudaStream_t s = h->stream;
  const uint32_t tiles = (uint32_t)((n + kEvTile - 1) / kEvTile);
  uint32_t* hp = h->h_pinned;
  const int wide = h->num_sms * 8;

  // ---- E1: one pass over the events.  Its targets are sized by upper bounds (G, C <= n; signal ids < S_cap) and live in the
  // staging allocation next to the event copy, so nothing has to be counted first.  Walker streams number their signals
  // densely (runtime.rs:120-125), so S_cap = n + 2^20 holds them; a stream with larger ids reruns the pass with its exact bound.
  uint64_t S_cap = std::min<uint64_t>(n + (1u << 20), 0x7FFFFFFEull);
  uint32_t *es = nullptr, *sig_t = nullptr, *gate_t = nullptr, *conn_t = nullptr, *conn_sb = nullptr;
  uint2 *sig_meta = nullptr, *conn = nullptr;
  uint4* egates = nullptr;
  uint8_t* outmark = nullptr;
  unsigned long long* nid_state = nullptr;
  uint32_t *side_best = nullptr, *side_parent = nullptr, *side_nidf = nullptr, *side_eff = nullptr, *side_effp = nullptr;
  const uint4* d_ev = nullptr;
  uint64_t G = 0, C = 0, n_sig = 0;
  uint32_t S = 0, flags = 0;
  bool copied = false;
  constexpr int kEarlyAt = 40;  // word offset of the early totals in h_emit_status (behind the final status block)
  const bool early = defer && pk_dense && h->h_emit_status;
  bool chunked = false;            // the payload words arrive on the copy stream while the scatter already runs
  bool scatter_in_flight = false;  // early totals were used: the scatter (which may read the CALLER's device arrays) is still running
  auto settle = [&]() { if (scatter_in_flight) { cudaStreamSynchronize(s); scatter_in_flight = false; } };  // before any error return
  for (int attempt = 0; attempt < 2; ++attempt) {
    const size_t pk_kbytes = pk ? align256(n + 4) : 0;  // packed staging: kinds, then words
    const size_t ev_copy = pk ? (src.pk_on_device ? 0 : pk_kbytes + align256(4 * pk->n_words + 4)) : (ev_dev ? 0 : align256(16 * n));
    // Dense packed stream: S = n - G - C and n_words = 3G + 2C give S <= n - n_words/3 and C <= n_words/2 BEFORE anything is
    // counted, so the union-find arrays can be carved (here, in the staging allocation) and initialised on the side stream
    // while the count / scatter kernels run; the main stream joins just before the first Boruvka kernel.
    // (with implicit operands a connection may carry a single word: C <= n_words; S <= n - n_words/3 holds either way)
    const uint64_t S_side = pk_dense ? n - (pk->n_words + 2) / 3 + 1 : 0;
    const uint64_t effw_side = pk_dense ? pk->n_words / (pk_impl ? 32 : 64) + 4 : 0;
    const size_t side_bytes = pk_dense ? 3 * align256(4 * S_side) + 2 * align256(4 * effw_side) : 0;
    const size_t ev_need = side_bytes + ev_copy + 3 * align256(4 * ((size_t)tiles + 2)) + align256(8 * ((size_t)scan_tiles((uint64_t)tiles + 1, kScanItems) + 1)) + align256(16 * ((size_t)scan_tiles((uint64_t)tiles + 1, kScanItems) + 1)) +
                           align256(8 * ((size_t)scan_tiles(n / 32 + 2, kScanItems) + 1)) + align256(4 * ES_COUNT) + 256 + align256(4 * S_cap) + align256(8 * S_cap) + align256(16 * n) + 3 * align256(4 * n) + align256(8 * n) + align256(S_cap);
    if (ev_need > h->ev_bytes) {
      if (h->ev_buf) { cudaStreamSynchronize(s); cudaFree(h->ev_buf); h->ev_buf = nullptr; h->ev_bytes = 0; }
      if (!cuda_ok(h, cudaMalloc(&h->ev_buf, ev_need + ev_need / 16), "cudaMalloc(event staging)")) { cudaGetLastError(); return C2A_ERR_NO_MEMORY; }
      h->ev_bytes = ev_need + ev_need / 16;
      copied = false;
    }
    char* q = h->ev_buf + ev_copy;
    auto take = [&](size_t bytes) { char* r = q; q += align256(bytes); return r; };
    const uint32_t ctiles = scan_tiles((uint64_t)tiles + 1, kScanItems);
    uint32_t* tile_g = (uint32_t*)take(4 * ((size_t)tiles + 2));
    uint32_t* tile_c = (uint32_t*)take(4 * ((size_t)tiles + 2));
    uint32_t* tile_i = (uint32_t*)take(4 * ((size_t)tiles + 2));
    unsigned long long* cnt_state = (unsigned long long*)take(16 * ((size_t)ctiles + 1));  // look-back states of the two count scans
    unsigned long long* cnt_state_i = (unsigned long long*)take(8 * ((size_t)ctiles + 1));  // ... and of the implicit-operand counts
    nid_state = (unsigned long long*)take(8 * ((size_t)scan_tiles(n / 32 + 2, kScanItems) + 1));  // ... and of the effective-connection bitmap scan (C <= n)
    es = (uint32_t*)take(4 * ES_COUNT);
    uint32_t* d_arrived = (uint32_t*)take(256);  // (outside the block the first memset clears: it is written by the copy stream)
    sig_t = (uint32_t*)take(4 * S_cap);
    sig_meta = (uint2*)take(8 * S_cap);
    egates = (uint4*)take(16 * n);
    gate_t = (uint32_t*)take(4 * n);
    conn_t = (uint32_t*)take(4 * n);
    conn_sb = (uint32_t*)take(4 * n);
    conn = (uint2*)take(8 * n);
    outmark = (uint8_t*)take(S_cap);
    if (pk_dense) {
      side_best = (uint32_t*)take(4 * S_side);
      side_parent = (uint32_t*)take(4 * S_side);
      side_nidf = (uint32_t*)take(4 * S_side);
      side_eff = (uint32_t*)take(4 * effw_side);
      side_effp = (uint32_t*)take(4 * effw_side);
      // (the side stream starts behind whatever an earlier call may have left running on the main stream - a call that returned an
      //  error before its final synchronisation - since these fills reuse the staging block those kernels work on)
      cudaEventRecord(h->ev_main, s);
      cudaStreamWaitEvent(h->stream2, h->ev_main, 0);
      cudaMemsetAsync(outmark, 0, std::min<uint64_t>(S_cap, n + 1), h->stream2);
      cudaEventRecord(h->ev_side2, h->stream2);
      cudaMemsetAsync(side_best, 0xFF, 4 * S_side, h->stream2);
      cudaMemsetAsync(side_nidf, 0, 4 * S_side, h->stream2);
      cudaMemsetAsync(side_eff, 0, 4 * effw_side, h->stream2);
      k_iota<<<grid_for(h, (const void*)k_iota, kBlock, S_side), kBlock, 0, h->stream2>>>(side_parent, (uint32_t)S_side);
      h->launches++;
      cudaEventRecord(h->ev_side, h->stream2);
    }
    d_ev = ev_dev ? (const uint4*)ev_dev : (const uint4*)h->ev_buf;
    const uint8_t* d_kinds = pk ? (src.pk_on_device ? pk->kinds : (const uint8_t*)h->ev_buf) : nullptr;
    const uint32_t* d_words = pk ? (src.pk_on_device ? pk->words : (const uint32_t*)(h->ev_buf + pk_kbytes)) : nullptr;

    if (!copied) {
      phase_begin(h, "h2d");
      if (pk) {
        if (n && !src.pk_on_device) {
          // A big dense stream in the one-call form: the count pass needs the kind bytes only, so the payload words follow on a copy
          // stream of their own, in chunks, each chunk followed by a 4-byte copy of the running word count - the scatter (whose tiles
          // use ascending word ranges) starts as soon as the counts are scanned and consumes the payload while it is still arriving.
          chunked = early && pk_impl && h->stream3 && pk->n_words >= (1u << 22) && pk->n_words < (1ull << 32);
          if (chunked) cudaMemsetAsync(d_arrived, 0, 4, s);
          if (!cuda_ok(h, cudaMemcpyAsync(h->ev_buf, pk->kinds, n, cudaMemcpyHostToDevice, s), "kinds H2D")) return C2A_ERR_CUDA;
          if (chunked) {
            cudaEventRecord(h->ev_main, s);
            cudaStreamWaitEvent(h->stream3, h->ev_main, 0);  // PCIe carries the kind bytes first
            uint32_t* cum = h->h_emit_status + 48;           // pinned: the running word counts, one per chunk
            constexpr int kChunks = 4;  // (every chunk costs two copy operations: 12 chunks lost to their latency what the finer overlap won)
            const uint64_t per = ((pk->n_words + kChunks - 1) / kChunks + 3) & ~3ull;
            int ci = 0;
            for (uint64_t off = 0; off < pk->n_words; off += per, ++ci) {
              const uint64_t len = std::min<uint64_t>(per, pk->n_words - off);
              cum[ci] = (uint32_t)(off + len);
              if (!cuda_ok(h, cudaMemcpyAsync(h->ev_buf + pk_kbytes + 4 * off, pk->words + off, 4 * len, cudaMemcpyHostToDevice, h->stream3), "words H2D (chunk)") ||
                  !cuda_ok(h, cudaMemcpyAsync(d_arrived, cum + ci, 4, cudaMemcpyHostToDevice, h->stream3), "words H2D (progress)")) {
                cudaStreamSynchronize(h->stream3);
                cudaStreamSynchronize(s);
                return C2A_ERR_CUDA;
              }
            }
            cudaEventRecord(h->ev_copy, h->stream3);
          } else if (pk->n_words && !cuda_ok(h, cudaMemcpyAsync(h->ev_buf + pk_kbytes, pk->words, 4 * pk->n_words, cudaMemcpyHostToDevice, s), "words H2D")) return C2A_ERR_CUDA;
        }
      } else if (n && !ev_dev && !cuda_ok(h, cudaMemcpyAsync(h->ev_buf, ev, 16 * n, cudaMemcpyHostToDevice, s), "events H2D")) return C2A_ERR_CUDA;
      phase_end(h);
      copied = true;
    }
    phase_begin(h, "init");
    cudaMemsetAsync(cnt_state, 0, (char*)es + 4 * ES_COUNT - (char*)cnt_state, s);  // look-back states of the count scans + the status block (carved back to back)
    if (!pk_dense) {
      cudaMemsetAsync(sig_t, 0xFF, 4 * S_cap, s);
      cudaMemsetAsync(outmark, 0, S_cap, s);
    }  // (dense: outmark is cleared on the side stream, beside the count pass - the scatter waits for it)
    phase_end(h);
    if (tiles) {
      const uint32_t egrid = std::min<uint32_t>(tiles, (uint32_t)wide);
      phase_begin(h, "k_ev_count");
      if (pk) LAUNCH(h, k_pk_count, egrid, kBlock, d_kinds, n, tiles, pk_impl ? 1u : 0u, tile_g, tile_c, tile_i, es);
      else LAUNCH(h, k_ev_count, std::min<uint32_t>(tiles, (uint32_t)grid_for(h, (const void*)k_ev_count, kBlock, n)), kBlock, d_ev, n, tiles, tile_g, tile_c, es);
      phase_end(h);
      phase_begin(h, "k_scan_u32");
      ScanJobs jobs{{tile_g, tile_c, tile_i}, {cnt_state, cnt_state + ctiles + 1, cnt_state_i}};
      LAUNCH(h, k_scan_u32_multi, dim3(scan_tiles(tiles, kScanItems), pk_impl ? 3 : 2), kBlock, jobs, tiles);
      phase_end(h);
      if (early) {
        // The host needs the totals only to size what FOLLOWS the scatter, and they exist once the count scans are done: the side
        // stream copies them out while the scatter runs, the host waits for that copy instead of for the scatter (whose own flags
        // are read with the final status - a deferred emit looks at them after the build's synchronisation anyway).
        uint32_t* e = h->h_emit_status + kEarlyAt;
        cudaEventRecord(h->ev_main, s);
        cudaStreamWaitEvent(h->stream2, h->ev_main, 0);
        cudaMemcpyAsync(e + 0, tile_g + tiles, 4, cudaMemcpyDeviceToHost, h->stream2);
        cudaMemcpyAsync(e + 1, tile_c + tiles, 4, cudaMemcpyDeviceToHost, h->stream2);
        if (pk_impl) cudaMemcpyAsync(e + 2, tile_i + tiles, 4, cudaMemcpyDeviceToHost, h->stream2);
        cudaMemcpyAsync(e + 3, es + ES_FLAGS, 4, cudaMemcpyDeviceToHost, h->stream2);  // the count pass's verdict on kinds and ops
        cudaEventRecord(h->ev_counts, h->stream2);
      }
      if (pk_dense) cudaStreamWaitEvent(s, h->ev_side2, 0);  // outmark is clear
      phase_begin(h, "k_ev_scatter");
      // persistent CTAs: exactly one resident wave (a partial second wave would run on a fraction of the SMs)
      if (pk && pk_impl) {
        if (chunked) LAUNCH(h, (k_pk_scatter_t<2, true>), std::min<uint32_t>(tiles, (uint32_t)grid_for(h, (const void*)k_pk_scatter_t<2, true>, kPkBlock, n)), kPkBlock, d_kinds, d_words, n, pk->n_words, 1u, tiles, (uint32_t)S_cap, tile_g, tile_c, (const uint32_t*)tile_i, sig_t, sig_meta,
               egates, gate_t, conn, conn_t, conn_sb, outmark, es, (const uint32_t*)d_arrived);
        else LAUNCH(h, k_pk_scatter_t<2>, std::min<uint32_t>(tiles, (uint32_t)grid_for(h, (const void*)k_pk_scatter_t<2>, kPkBlock, n)), kPkBlock, d_kinds, d_words, n, pk->n_words, 1u, tiles, (uint32_t)S_cap, tile_g, tile_c, (const uint32_t*)tile_i, sig_t, sig_meta,
               egates, gate_t, conn, conn_t, conn_sb, outmark, es, (const uint32_t*)nullptr);
        if (chunked) LAUNCH(h, (k_pk_scatter_t<1, true>), std::min<uint32_t>(tiles, (uint32_t)grid_for(h, (const void*)k_pk_scatter_t<1, true>, kPkBlock, n)), kPkBlock, d_kinds, d_words, n, pk->n_words, 1u, tiles, (uint32_t)S_cap, tile_g, tile_c, (const uint32_t*)tile_i, sig_t, sig_meta,
               egates, gate_t, conn, conn_t, conn_sb, outmark, es, (const uint32_t*)d_arrived);
        else LAUNCH(h, k_pk_scatter_t<1>, std::min<uint32_t>(tiles, (uint32_t)grid_for(h, (const void*)k_pk_scatter_t<1>, kPkBlock, n)), kPkBlock, d_kinds, d_words, n, pk->n_words, 1u, tiles, (uint32_t)S_cap, tile_g, tile_c, (const uint32_t*)tile_i, sig_t, sig_meta,
               egates, gate_t, conn, conn_t, conn_sb, outmark, es, (const uint32_t*)nullptr);
      } else if (pk) LAUNCH(h, k_pk_scatter_t<0>, std::min<uint32_t>(tiles, (uint32_t)grid_for(h, (const void*)k_pk_scatter_t<0>, kPkBlock, n)), kPkBlock, d_kinds, d_words, n, pk->n_words, pk_dense ? 1u : 0u, tiles, (uint32_t)S_cap, tile_g, tile_c, (const uint32_t*)nullptr, sig_t, sig_meta,
                     egates, gate_t, conn, conn_t, conn_sb, outmark, es, (const uint32_t*)nullptr);
      else LAUNCH(h, k_ev_scatter, std::min<uint32_t>(tiles, (uint32_t)grid_for(h, (const void*)k_ev_scatter, kBlock, n)), kBlock, d_ev, n, tiles, (uint32_t)S_cap, tile_g, tile_c, sig_t, sig_meta, egates, gate_t, conn, conn_t, conn_sb, es);
      phase_end(h);
      if (chunked) cudaStreamWaitEvent(s, h->ev_copy, 0);  // (the main stream is ordered behind the copy stream again)
      cudaMemcpyAsync(es + ES_NGATE, tile_g + tiles, 4, cudaMemcpyDeviceToDevice, s);  // totals land behind the scanned arrays
      cudaMemcpyAsync(es + ES_NCONN, tile_c + tiles, 4, cudaMemcpyDeviceToDevice, s);
      if (pk_impl) cudaMemcpyAsync(es + ES_NIMPL, tile_i + tiles, 4, cudaMemcpyDeviceToDevice, s);
    }
    bool have_counts = false;
    uint64_t n_impl = 0;
    if (early && tiles) {
      if (!cuda_ok(h, cudaEventSynchronize(h->ev_counts), "count sync")) return C2A_ERR_CUDA;
      const uint32_t* e = h->h_emit_status + kEarlyAt;
      G = e[0];
      C = e[1];
      n_impl = pk_impl ? e[2] : 0;
      // anything irregular goes the ordinary way: wait for the scatter and look at the whole status block
      have_counts = e[3] == 0 && G + C <= n && pk->n_words == 3 * G + 2 * C - n_impl;
      if (have_counts) { n_sig = n - G - C; S = (uint32_t)n_sig; flags = 0; scatter_in_flight = true; }  // dense ids: the table bound is the number of declarations
    }
    if (!have_counts) {
      cudaMemcpyAsync(hp, es, 4 * ES_COUNT, cudaMemcpyDeviceToHost, s);
      if (!cuda_ok(h, cudaStreamSynchronize(s), "scatter sync")) return C2A_ERR_CUDA;
      if (!cuda_ok(h, cudaGetLastError(), "event scatter")) return C2A_ERR_CUDA;
      G = hp[ES_NGATE];
      C = hp[ES_NCONN];
      S = hp[ES_SBOUND];
      flags = hp[ES_FLAGS];
      n_sig = n - G - C;
      n_impl = pk_impl ? hp[ES_NIMPL] : 0;
    }
    if (pk && !(flags & EF_BAD_KIND) && pk->n_words != 3 * G + 2 * C - n_impl + (pk_dense ? 0 : n_sig))
      return fail(h, C2A_ERR_INVALID_ARGUMENT, "n_words (%llu) does not match the kinds (%llu gates, %llu connections, %llu signals)",
                  (unsigned long long)pk->n_words, (unsigned long long)G, (unsigned long long)C, (unsigned long long)n_sig);
    if (!(flags & (EF_BAD_KIND | EF_SPARSE)) && (uint64_t)S > 4 * n_sig + (1u << 20)) flags |= EF_SPARSE;  // a dense table would be mostly holes
    if ((flags & EF_CAP) && !(flags & (EF_BAD_KIND | EF_BAD_OP | EF_SPARSE)) && attempt == 0) {
      S_cap = S;  // valid but gappy ids: once more with the exact table bound
      continue;
    }
    break;
  }
  flags &= ~(uint32_t)EF_CAP;
  if (info) { info->n_gates = G; info->n_connections = C; info->n_signals = n_sig; info->signal_bound = S; }

  std::vector<c2a_event> ev_back;
  bool prod1_pending = false;  // a memset of the slab is in flight on the side stream
  auto decline = [&](uint32_t f) -> int {
    if (info) info->decline_flags = f;
    if (prod1_pending) { cudaStreamWaitEvent(s, h->ev_side, 0); prod1_pending = false; }  // the host path reuses the slab
    if (!ev && n) {  // the exact replay needs AoS events on the host
      ev_back.resize(n);
      if (pk) {
        c2a_packed_events hp_pk = *pk;
        std::vector<uint8_t> kb;
        std::vector<uint32_t> wb;
        if (src.pk_on_device) {
          kb.resize(n);
          wb.resize(pk->n_words);
          if (!cuda_ok(h, cudaMemcpy(kb.data(), pk->kinds, n, cudaMemcpyDeviceToHost), "kinds D2H for the host replay")) return C2A_ERR_CUDA;
          if (pk->n_words && !cuda_ok(h, cudaMemcpy(wb.data(), pk->words, 4 * pk->n_words, cudaMemcpyDeviceToHost), "words D2H for the host replay")) return C2A_ERR_CUDA;
          hp_pk.kinds = kb.data();
          hp_pk.words = wb.data();
        }
        if (c2a_unpack_events(&hp_pk, ev_back.data()) != C2A_OK) return fail(h, C2A_ERR_INVALID_ARGUMENT, "n_words does not match the kinds");
      } else if (!cuda_ok(h, cudaMemcpy(ev_back.data(), ev_dev, 16 * n, cudaMemcpyDeviceToHost), "events D2H for the host replay")) return C2A_ERR_CUDA;
      ev = ev_back.data();
    }
    int r = emit_host_path(h, ev, n, S, (f & EF_SPARSE) != 0, info, err_event);
    if (info) info->decline_flags = f;
    phases_collect(h);
    return r;
  };
  if (flags) return decline(flags);

  // ---- exact sizes are known: one slab for the resident result, the emit scratch and the build that follows
  const uint32_t NB_ub = (uint32_t)std::min<uint64_t>(n_sig + C + 1, 0x7FFFFFFEull);
  BuildPlan bp{G, NB_ub, 0, 0, true};
  size_t resident = align256(16 * G) + align256(4 * (size_t)S) + align256(4 * (size_t)NB_ub);
  size_t build_need = core_scratch_bytes(bp, 1u << 20) + align256(16 * G) + align256(4 * G) + align256(4 * (size_t)NB_ub);
  slab_reset(h);
  emit_drop_host(h);
  if (!slab_reserve(h, resident + std::max(emit_scratch_bytes(G, C, S), build_need))) { settle(); return C2A_ERR_NO_MEMORY; }
  uint4* d_gates = (uint4*)slab_alloc(h, 16 * G);
  uint32_t* nos = (uint32_t*)slab_alloc(h, 4 * (size_t)S);
  uint32_t* prod1 = (uint32_t*)slab_alloc(h, 4 * (size_t)NB_ub);  // producer map of the build (K1), filled by k_ev_gates
  if (!prod1) { settle(); return fail(h, C2A_ERR_NO_MEMORY, "scratch slab exhausted"); }
  const size_t keep = h->slab_used;
  const uint32_t effw = (uint32_t)(C / 32 + 1);  // bitmap words; one spare bit at least, so rank(C) = total is addressable
  uint32_t* parent = pk_dense ? side_parent : (uint32_t*)slab_alloc(h, 4 * (size_t)S);
  uint32_t* best = pk_dense ? side_best : (uint32_t*)slab_alloc(h, 4 * (size_t)S);
  uint32_t* nidf = pk_dense ? side_nidf : (uint32_t*)slab_alloc(h, 4 * (size_t)S);
  uint32_t* eff = pk_dense ? side_eff : (uint32_t*)slab_alloc(h, 4 * ((size_t)effw + 3));
  uint32_t* effp = pk_dense ? side_effp : (uint32_t*)slab_alloc(h, 4 * ((size_t)effw + 3));
  if (!parent || !best || !nidf || !eff || !effp) { settle(); return fail(h, C2A_ERR_NO_MEMORY, "scratch slab exhausted"); }
  uint32_t* cur = (uint32_t*)slab_alloc(h, 4 * C);
  uint4* cand = (uint4*)slab_alloc(h, 16 * C);
  if (!cand) { settle(); return fail(h, C2A_ERR_NO_MEMORY, "scratch slab exhausted"); }
  unsigned long long* tile_state = nid_state;  // in the staging block: zeroed by the first memset of the call

  phase_begin(h, "init");
  if (pk_dense) cudaStreamWaitEvent(s, h->ev_side, 0);  // initialised on the side stream while E0 / E1 ran
  else {
    cudaMemsetAsync(best, 0xFF, 4 * (size_t)S, s);
    cudaMemsetAsync(eff, 0, 4 * ((size_t)effw + 3), s);
    if (S) LAUNCH(h, k_iota, grid_for(h, (const void*)k_iota, kBlock, S), kBlock, parent, S);
  }
  phase_end(h);
  // prod1 is zeroed on the side stream next to the Boruvka / node-id kernels; k_ev_gates joins.  (Nothing on the main stream
  // touches the slab at this point - it is idle after the scatter sync, or still running the scatter, which writes the staging
  // block only - so the side stream needs no event from it.)
  cudaMemsetAsync(prod1, 0, 4 * (size_t)NB_ub, h->stream2);
  cudaEventRecord(h->ev_side, h->stream2);
  prod1_pending = true;
  // E2 runs only when the ids are explicit; a dense stream was validated inside the scatter
  phase_begin(h, "k_ev_check_gates");
  if (G && !pk_dense) LAUNCH(h, k_ev_check_gates, grid_for(h, (const void*)k_ev_check_gates, kBlock, G), kBlock, egates, gate_t, (uint32_t)G, S, sig_t, outmark, es);
  phase_end(h);
  phase_begin(h, "k_ev_check_conns");
  if (C && !pk_dense) LAUNCH(h, k_ev_check_conns, grid_for(h, (const void*)k_ev_check_conns, kBlock, C), kBlock, conn, conn_t, (uint32_t)C, S, sig_t, es);
  phase_end(h);

  // ---- Boruvka rounds.  Round r: candidates counted in *ncand, surviving (undecided) edges in *ncur.
  uint32_t rounds_issued = 0;
  auto msf_round = [&](uint32_t* ncur_prev, uint32_t* ncand, uint32_t* ncur) {
    uint32_t tag = (6u - (rounds_issued % 7u)) << 29;
    if (rounds_issued && (rounds_issued % 7u) == 0) cudaMemsetAsync(best, 0xFF, 4 * (size_t)S, s);  // tags wrapped: forget the old minima
    phase_begin(h, "k_msf_pick");
    const bool first = rounds_issued == 0;
    if (first) LAUNCH(h, k_msf_pick_first, grid_for(h, (const void*)k_msf_pick_first, kBlock, C), kBlock, conn, (uint32_t)C, best, tag, ncand);
    else LAUNCH(h, k_msf_pick, wide, kBlock, conn, cur, ncur_prev, parent, best, tag, cand, ncand);
    phase_end(h);
    phase_begin(h, "k_msf_hook");
    if (first) LAUNCH(h, k_msf_hook_t<true>, wide, kBlock, cand, conn, ncand, parent, best, tag, eff, cur, ncur, es + ES_ROUNDS);
    else LAUNCH(h, k_msf_hook_t<false>, wide, kBlock, cand, conn, ncand, parent, best, tag, eff, cur, ncur, es + ES_ROUNDS);
    phase_end(h);
    ++rounds_issued;
  };
  // ---- node ids (re-issued when the speculative rounds turn out not to have finished the forest)
  int nid_runs = 0;
  auto node_ids = [&](bool sync = true) {
    uint32_t stiles = scan_tiles(C + 1, kScanItems);
    phase_begin(h, "init");
    if (nid_runs) {  // a second run (the forest needed more rounds): forget the first one
      cudaMemsetAsync(tile_state, 0, 8 * (size_t)stiles, s);
      cudaMemsetAsync(es + ES_NDECL, 0, 4, s);
    }
    if (!pk_dense || nid_runs) cudaMemsetAsync(nidf, 0, 4 * (size_t)S, s);  // (dense, first run: zeroed on the side stream)
    ++nid_runs;
    phase_end(h);
    // effp = exclusive scan of popcount(eff words); effp[effw] receives the total (= effective connections)
    phase_begin(h, "k_scan_u32");
    LAUNCH(h, k_scan_u32_t<true>, scan_tiles(effw, kScanItems), kBlock, eff, effp, effw, tile_state, (uint32_t*)nullptr, (const uint32_t*)nullptr, 0);
    phase_end(h);
    phase_begin(h, "k_ev_nid_edges");
    if (C) LAUNCH(h, k_ev_nid_edges, grid_for(h, (const void*)k_ev_nid_edges, kBlock, C), kBlock, (uint32_t)C, conn, conn_sb, eff, effp, parent, nidf);
    phase_end(h);
    phase_begin(h, "k_ev_finalize");
    if (S) LAUNCH(h, k_ev_finalize, grid_for(h, (const void*)k_ev_finalize, kBlock, S), kBlock, S, pk_dense ? (const uint32_t*)nullptr : sig_t, sig_meta, outmark, eff, effp, parent, nidf, nos, es);
    phase_end(h);
    if (prod1_pending) { cudaStreamWaitEvent(s, h->ev_side, 0); prod1_pending = false; }
    else cudaMemsetAsync(prod1, 0, 4 * (size_t)NB_ub, s);  // second run (the forest needed more rounds): forget the first run's producers
    phase_begin(h, "k_ev_gates");
    if (G) LAUNCH(h, k_ev_gates, grid_for(h, (const void*)k_ev_gates, kBlock, G), kBlock, egates, (uint32_t)G, S, nos, d_gates, prod1, NB_ub);
    phase_end(h);
    uint32_t* dst = sync ? hp : h->h_emit_status;  // (deferred: a buffer of its own - the build re-stages, even re-allocates, hp)
    // deferred: the copies ride on the side stream (the build joins it in front of its wire kernels, so its final synchronisation
    // covers them) - the main stream goes straight on to the build
    cudaStream_t cs = s;
    if (!sync) {
      cudaEventRecord(h->ev_main, s);
      cudaStreamWaitEvent(h->stream2, h->ev_main, 0);
      cs = h->stream2;
    }
    cudaMemcpyAsync(dst, es, 4 * ES_COUNT, cudaMemcpyDeviceToHost, cs);
    cudaMemcpyAsync(dst + ES_COUNT, effp + effw, 4, cudaMemcpyDeviceToHost, cs);
    if (!sync) return true;
    if (!cuda_ok(h, cudaStreamSynchronize(s), "emit sync")) return false;
    return cuda_ok(h, cudaGetLastError(), "emit kernels");
  };

  if (C) {
    for (int r = 0; r < kSpecMsf; ++r)
      msf_round(r ? es + ES_MC0 + 2 * (r - 1) + 1 : nullptr, es + ES_MC0 + 2 * r, es + ES_MC0 + 2 * r + 1);
  }
  if (defer && pk_dense && h->h_emit_status) {
    // provisional result: exact gates / signals, node ids bounded by signals + connections; fixed up by the caller after the build's sync
    node_ids(false);
    uint64_t nc_ub = n_sig + C;
    if (defer->wire_cap) nc_ub = std::min<uint64_t>(nc_ub, defer->wire_cap - 1);
    defer->pending = true;
    defer->es = es;
    defer->n_sig = n_sig;
    defer->C = C;
    h->slab_used = keep;
    h->slab_keep = keep;
    h->emitted.valid = true;
    h->emitted.nos_valid = true;
    h->emitted.gates_off = (char*)d_gates - h->slab;
    h->emitted.nos_off = (char*)nos - h->slab;
    h->emitted.prod1_valid = true;
    h->emitted.prod1_off = (char*)prod1 - h->slab;
    h->emitted.G = G;
    h->emitted.node_count = (uint32_t)nc_ub;
    h->emitted.signal_bound = S;
    h->emitted.wire = nullptr;
    return C2A_OK;
  }
  if (!node_ids()) return C2A_ERR_CUDA;
  if (C && hp[ES_MC0 + 2 * (kSpecMsf - 1)] != 0 && hp[ES_MC0 + 2 * (kSpecMsf - 1) + 1] != 0) {
    // the last speculative round still had candidates and left edges undecided: finish with host-checked rounds, then redo the ids.
    // (eff[] is only ever added to, parent[] only compressed/hooked further: the re-run starts from a consistent state.)
    cudaMemcpyAsync(es + ES_NCUR, es + ES_MC0 + 2 * (kSpecMsf - 1) + 1, 4, cudaMemcpyDeviceToDevice, s);
    while (true) {
      cudaMemcpyAsync(es + ES_PREV, es + ES_NCUR, 4, cudaMemcpyDeviceToDevice, s);  // the previous round's live count
      cudaMemsetAsync(es + ES_NCUR, 0, 8, s);                                         // ES_NCUR, ES_NCAND
      msf_round(es + ES_PREV, es + ES_NCAND, es + ES_NCUR);
      cudaMemcpyAsync(hp, es + ES_NCUR, 8, cudaMemcpyDeviceToHost, s);
      if (!cuda_ok(h, cudaStreamSynchronize(s), "msf sync")) return C2A_ERR_CUDA;
      if (hp[1] == 0 || hp[0] == 0) break;  // no candidate edges at all, or none left undecided
      if (rounds_issued > 96) return fail(h, C2A_ERR_CUDA, "Boruvka did not converge");
    }
    if (!node_ids()) return C2A_ERR_CUDA;
  }
  flags = hp[ES_FLAGS];
  if (hp[ES_NDECL] != n_sig) flags |= EF_DUPLICATE;  // fewer distinct ids than signal events
  if (flags) return decline(flags);
  const uint32_t n_eff = C ? hp[ES_COUNT] : 0u;

  h->slab_used = keep;
  h->slab_keep = keep;
  h->emitted.valid = true;
  h->emitted.nos_valid = true;
  h->emitted.gates_off = (char*)d_gates - h->slab;
  h->emitted.nos_off = (char*)nos - h->slab;
  h->emitted.prod1_valid = true;
  h->emitted.prod1_off = (char*)prod1 - h->slab;
  h->emitted.G = G;
  h->emitted.node_count = (uint32_t)(n_sig + n_eff);
  h->emitted.signal_bound = S;
  h->emitted.wire = nullptr;
  if (info) {
    info->n_effective = n_eff;
    info->node_count = h->emitted.node_count;
    info->path = C2A_EMIT_PATH_DEVICE;
    info->rounds = hp[ES_ROUNDS];
  }
  phases_collect(h);
  return C2A_OK;
}

int c2a_emit_events_device(c2a_handle* h, const c2a_event* ev, uint64_t n, c2a_emit_info* info, uint64_t* err_event) {
  EmitSrc src;
  src.ev_host = ev;
  return emit_events_impl(h, src, n, info, err_event);
}
int c2a_emit_events_resident(c2a_handle* h, const c2a_event* d_ev, uint64_t n, c2a_emit_info* info, uint64_t* err_event) {
  if (n && !d_ev) return fail(h, C2A_ERR_INVALID_ARGUMENT, "null event array");
  EmitSrc src;
  src.ev_dev = d_ev;
  return emit_events_impl(h, src, n, info, err_event);
}
int c2a_emit_packed_device(c2a_handle* h, const c2a_packed_events* pk, c2a_emit_info* info, uint64_t* err_event) {
  if (!pk) return fail(h, C2A_ERR_INVALID_ARGUMENT, "null packed stream");
  EmitSrc src;
  src.pk = pk;
  return emit_events_impl(h, src, pk->n_events, info, err_event);
}
int c2a_emit_packed_resident(c2a_handle* h, const c2a_packed_events* d_pk, c2a_emit_info* info, uint64_t* err_event) {
  if (!d_pk) return fail(h, C2A_ERR_INVALID_ARGUMENT, "null packed stream");
  EmitSrc src;
  src.pk = d_pk;
  src.pk_on_device = true;
  return emit_events_impl(h, src, d_pk->n_events, info, err_event);
}

// Compressed stream: only the literal ranges and the replay records cross PCIe; the replayed instances are expanded in HBM,
// generation by generation, and the expanded stream goes through the resident packed path.
int c2a_emit_compressed_device(c2a_handle* h, const c2a_compressed_events* cx, c2a_emit_info* info, uint64_t* err_event) {
  if (!h || !cx) return C2A_ERR_INVALID_ARGUMENT;
  const uint64_t n = cx->n_events, nw = cx->n_words, nr = cx->n_replays;
  int st = check_sizes(h, n, 1);
  if (st) return st;
  if ((n && !cx->kinds) || (nw && !cx->words) || (nr && !cx->replays)) return fail(h, C2A_ERR_INVALID_ARGUMENT, "null compressed arrays");
  if (nw > 3 * n) return fail(h, C2A_ERR_INVALID_ARGUMENT, "n_words does not match the kinds");
  // ---- records: ascending, disjoint, inside the arrays, sources before destinations, generations consistent with the order
  uint64_t k_at = 0, w_at = 0, lit_k = 0, lit_w = 0;
  uint32_t max_gen = 0;
  for (uint64_t i = 0; i < nr; ++i) {
    const c2a_replay& r = cx->replays[i];
    if (r.k_dst < k_at || r.w_dst < w_at || r.k_dst > n || r.w_dst > nw || r.k_len > n - r.k_dst || r.w_len > nw - r.w_dst ||
        r.k_src > r.k_dst || r.w_src > r.w_dst || r.k_src + r.k_len > r.k_dst ||
        r.w_src + r.w_len > r.w_dst || r.gen == 0 || r.gen > 1u << 20)
      return fail(h, C2A_ERR_INVALID_ARGUMENT, "replay record %llu is inconsistent", (unsigned long long)i);
    lit_k += r.k_dst - k_at;
    lit_w += r.w_dst - w_at;
    k_at = r.k_dst + r.k_len;
    w_at = r.w_dst + r.w_len;
    max_gen = std::max(max_gen, r.gen);
  }
  lit_k += n - k_at;
  lit_w += nw - w_at;
  // generations: a record may only read what is complete when its launch starts - literal bytes, or the destination of a record of a
  // SMALLER generation.  (Records ascend by destination: the destinations that intersect a source range are a run of records,
  // found by binary search; the largest generation in the run by a sparse table.)  Without this an inconsistent record would read
  // bytes that are unwritten in this launch, or stale ones from the previous call, and expand to a plausible but wrong stream.
  if (nr) {
    std::vector<std::vector<uint32_t>> tab;
    if (max_gen > 1) {  // (one generation: any destination inside a source range is already an error - no table needed)
      tab.emplace_back(nr);
      for (uint64_t i = 0; i < nr; ++i) tab[0][i] = cx->replays[i].gen;
      for (uint64_t len = 2, lv = 1; len <= nr; len <<= 1, ++lv) {
        tab.emplace_back(nr - len + 1);
        for (uint64_t i = 0; i + len <= nr; ++i) tab[lv][i] = std::max(tab[lv - 1][i], tab[lv - 1][i + len / 2]);
      }
    }
    auto range_max = [&](uint64_t lo, uint64_t hi) -> uint32_t {  // max gen of records [lo, hi)
      if (lo >= hi) return 0;
      if (max_gen <= 1) return 1;
      const int lv = 63 - __builtin_clzll(hi - lo);
      return std::max(tab[lv][lo], tab[lv][hi - (1ull << lv)]);
    };
    for (uint64_t i = 0; i < nr; ++i) {
      const c2a_replay& r = cx->replays[i];
      for (int which = 0; which < 2; ++which) {
        const uint64_t src = which ? r.w_src : r.k_src, len = which ? r.w_len : r.k_len;
        if (!len) continue;
        auto dst_of = [&](uint64_t j) { return which ? cx->replays[j].w_dst : cx->replays[j].k_dst; };
        auto len_of = [&](uint64_t j) { return which ? cx->replays[j].w_len : cx->replays[j].k_len; };
        // first record whose destination ends behind src, first record whose destination starts at or behind src + len (both < i)
        uint64_t lo = 0, hi = i;
        while (lo < hi) { uint64_t m = (lo + hi) / 2; if (dst_of(m) + len_of(m) > src) hi = m; else lo = m + 1; }
        const uint64_t first = lo;
        lo = first; hi = i;
        while (lo < hi) { uint64_t m = (lo + hi) / 2; if (dst_of(m) >= src + len) hi = m; else lo = m + 1; }
        if (range_max(first, lo) >= r.gen)
          return fail(h, C2A_ERR_INVALID_ARGUMENT, "replay record %llu reads the destination of a record of generation >= its own (%u)", (unsigned long long)i, r.gen);
      }
    }
  }
  cudaStream_t s = h->stream;
  // ---- chunk tables: generation 0 = the literal ranges (source: the packed literal staging), generation g = records of gen g
  auto n_chunks_of = [](uint64_t len) { return (len + kCxChunk - 1) / kCxChunk; };
  std::vector<std::vector<CxChunk>> kch(max_gen + 1), wch(max_gen + 1);
  if (max_gen == 1) { kch[1].reserve(nr + n / kCxChunk + 1); wch[1].reserve(nr + nw / kCxChunk + 1); }
  auto add_chunks = [&](std::vector<CxChunk>& v, uint64_t dst, uint64_t src, uint64_t len, uint32_t delta) {
    for (uint64_t o = 0; o < len; o += kCxChunk) v.push_back(CxChunk{dst + o, src + o, (uint32_t)std::min<uint64_t>(kCxChunk, len - o), delta});
  };
  const size_t lit_k_bytes = align256(lit_k + 16), lit_w_bytes = align256(4 * lit_w + 16);
  size_t table_entries = n_chunks_of(lit_k) + n_chunks_of(lit_w) + 2 * (nr + 2);
  for (uint64_t i = 0; i < nr; ++i) table_entries += n_chunks_of(cx->replays[i].k_len) + n_chunks_of(cx->replays[i].w_len);
  const size_t stage_need = lit_k_bytes + lit_w_bytes + align256(sizeof(CxChunk) * table_entries);
  if (stage_need > h->cx_pinned_bytes) {
    if (h->cx_pinned) { cudaStreamSynchronize(s); cudaFreeHost(h->cx_pinned); h->cx_pinned = nullptr; h->cx_pinned_bytes = 0; }
    if (!cuda_ok(h, cudaHostAlloc((void**)&h->cx_pinned, stage_need + stage_need / 8, cudaHostAllocDefault), "cudaHostAlloc(compressed staging)")) return C2A_ERR_NO_MEMORY;
    h->cx_pinned_bytes = stage_need + stage_need / 8;
  } else cudaStreamSynchronize(s);  // the staging may still be the source of the previous call's copy
  uint8_t* st_k = (uint8_t*)h->cx_pinned;
  uint32_t* st_w = (uint32_t*)(h->cx_pinned + lit_k_bytes);
  CxChunk* st_t = (CxChunk*)(h->cx_pinned + lit_k_bytes + lit_w_bytes);
  {  // pack the literal ranges, build the tables
    uint64_t pk = 0, pw = 0;
    k_at = w_at = 0;
    auto literal = [&](uint64_t k_end, uint64_t w_end) {
      if (k_end > k_at) { memcpy(st_k + pk, cx->kinds + k_at, k_end - k_at); add_chunks(kch[0], k_at, pk, k_end - k_at, 0); pk += k_end - k_at; }
      if (w_end > w_at) { memcpy(st_w + pw, cx->words + w_at, 4 * (w_end - w_at)); add_chunks(wch[0], w_at, pw, w_end - w_at, 0); pw += w_end - w_at; }
    };
    for (uint64_t i = 0; i < nr; ++i) {
      const c2a_replay& r = cx->replays[i];
      literal(r.k_dst, r.w_dst);
      add_chunks(kch[r.gen], r.k_dst, r.k_src, r.k_len, 0);
      add_chunks(wch[r.gen], r.w_dst, r.w_src, r.w_len, r.delta);
      k_at = r.k_dst + r.k_len;
      w_at = r.w_dst + r.w_len;
    }
    literal(n, nw);
  }
  size_t t_at = 0;
  std::vector<size_t> k_off(max_gen + 2), w_off(max_gen + 2);
  for (uint32_t g = 0; g <= max_gen; ++g) {
    k_off[g] = t_at; if (!kch[g].empty()) memcpy(st_t + t_at, kch[g].data(), sizeof(CxChunk) * kch[g].size()); t_at += kch[g].size();
    w_off[g] = t_at; if (!wch[g].empty()) memcpy(st_t + t_at, wch[g].data(), sizeof(CxChunk) * wch[g].size()); t_at += wch[g].size();
  }
  // ---- device: [kinds | words | literal kinds | literal words | tables]
  const size_t d_k_bytes = align256(n + 16), d_w_bytes = align256(4 * nw + 16), d_t_bytes = align256(sizeof(CxChunk) * (t_at + 1));
  const size_t dev_need = d_k_bytes + d_w_bytes + lit_k_bytes + lit_w_bytes + d_t_bytes;
  if (dev_need > h->cx_bytes) {
    if (h->cx_buf) { cudaStreamSynchronize(s); cudaFree(h->cx_buf); h->cx_buf = nullptr; h->cx_bytes = 0; }
    if (!cuda_ok(h, cudaMalloc(&h->cx_buf, dev_need + dev_need / 16), "cudaMalloc(compressed stream)")) { cudaGetLastError(); return C2A_ERR_NO_MEMORY; }
    h->cx_bytes = dev_need + dev_need / 16;
  }
  uint8_t* d_kinds = (uint8_t*)h->cx_buf;
  uint32_t* d_words = (uint32_t*)(h->cx_buf + d_k_bytes);
  uint8_t* d_lit_k = (uint8_t*)(h->cx_buf + d_k_bytes + d_w_bytes);
  uint32_t* d_lit_w = (uint32_t*)(h->cx_buf + d_k_bytes + d_w_bytes + lit_k_bytes);
  CxChunk* d_t = (CxChunk*)(h->cx_buf + d_k_bytes + d_w_bytes + lit_k_bytes + lit_w_bytes);
  phases_clear(h);
  if (!cuda_ok(h, cudaMemcpyAsync(d_lit_k, h->cx_pinned, lit_k_bytes + lit_w_bytes + sizeof(CxChunk) * t_at, cudaMemcpyHostToDevice, s), "compressed H2D")) return C2A_ERR_CUDA;
  const int wide = h->num_sms * 8;
  phase_begin(h, "k_cx_copy");
  for (uint32_t g = 0; g <= max_gen; ++g) {
    const uint32_t nk = (uint32_t)kch[g].size(), nwc = (uint32_t)wch[g].size();
    if (nk) LAUNCH(h, k_cx_copy_u8, std::min<uint32_t>(nk, (uint32_t)wide), kBlock, d_t + k_off[g], nk, g ? (const uint8_t*)d_kinds : (const uint8_t*)d_lit_k, d_kinds);
    if (nwc) LAUNCH(h, k_cx_copy_u32, std::min<uint32_t>(nwc, (uint32_t)wide), kBlock, d_t + w_off[g], nwc, g ? (const uint32_t*)d_words : (const uint32_t*)d_lit_w, d_words);
  }
  phase_end(h);
  if (!cuda_ok(h, cudaGetLastError(), "expand kernels")) return C2A_ERR_CUDA;
  c2a_packed_events pk{d_kinds, d_words, n, nw, cx->flags, 0};
  EmitSrc src;
  src.pk = &pk;
  src.pk_on_device = true;
  src.keep_phases = true;
  return emit_events_impl(h, src, n, info, err_event);
}

int c2a_emitted_fetch(c2a_handle* h, c2a_gate* gates_out, uint32_t* node_of_signal_out) {
  if (!h) return C2A_ERR_INVALID_ARGUMENT;
  if (!h->emitted.valid) return fail(h, C2A_ERR_INVALID_ARGUMENT, "no emitted circuit is resident on this handle");
  if (!cuda_ok(h, cudaSetDevice(h->device), "cudaSetDevice")) return C2A_ERR_CUDA;
  if (gates_out && h->emitted.G) cudaMemcpyAsync(gates_out, h->slab + h->emitted.gates_off, 16 * h->emitted.G, cudaMemcpyDeviceToHost, h->stream);
  if (node_of_signal_out) {
    if (!h->emitted.nos_valid) return fail(h, C2A_ERR_INVALID_ARGUMENT, "node_of_signal is not resident (sparse signal ids)");
    if (h->emitted.signal_bound) cudaMemcpyAsync(node_of_signal_out, h->slab + h->emitted.nos_off, 4 * (size_t)h->emitted.signal_bound, cudaMemcpyDeviceToHost, h->stream);
  }
  if (!cuda_ok(h, cudaStreamSynchronize(h->stream), "emitted fetch")) return C2A_ERR_CUDA;
  return C2A_OK;
}

// gate range [g_lo, g_hi) of the resident circuit (g_hi = ~0: all of it).  A proper sub-range is one shard of a sharded build
// (c2a_plan_shards_device): no dependency edge leaves it, so its producer map is rebuilt for the range alone.
static int emitted_build_impl(c2a_handle* h, const uint32_t* input_signals, uint32_t n_in, const uint32_t* output_signals, uint32_t n_out,
                              uint32_t* order_out, uint32_t* wire_of_node, c2a_gate* new_gates, uint32_t* wire_count, uint64_t* err_index,
                              bool outputs_on_device, uint64_t g_lo = 0, uint64_t g_hi = ~0ull, bool keep_phases = false,
                              const uint32_t* d_io_sigs_ready = nullptr /* the two lists, already on the device */, uint32_t* es_ready = nullptr /* a zeroed status block */) {
  if (!h) return C2A_ERR_INVALID_ARGUMENT;
  if (!h->emitted.valid) return fail(h, C2A_ERR_INVALID_ARGUMENT, "no emitted circuit is resident on this handle");
  const bool whole = g_lo == 0 && (g_hi == ~0ull || g_hi == h->emitted.G);
  if (!whole && (g_lo > g_hi || g_hi > h->emitted.G)) return fail(h, C2A_ERR_INVALID_ARGUMENT, "bad gate range");
  const uint64_t G = whole ? h->emitted.G : g_hi - g_lo;
  const uint32_t node_bound = h->emitted.node_count + 1;
  int st = check_sizes(h, G, node_bound);
  if (st) return st;
  if (!keep_phases) phases_clear(h);
  cudaStream_t s = h->stream;
  BuildPlan p{G, node_bound, n_in, n_out, true};
  const size_t n_pairs = (size_t)n_in + n_out;
  slab_reset_keep(h);
  size_t need = h->slab_keep + core_scratch_bytes(p, n_pairs) + align256(4 * n_pairs + 4) + align256(16 * G) + align256(4 * G) + align256(4 * (size_t)node_bound) + align256(4 * ES_COUNT);
  if (!slab_reserve(h, need)) return C2A_ERR_NO_MEMORY;  // (the slab may move: pointers into it are taken below)
  const uint4* d_gates = (const uint4*)(h->slab + h->emitted.gates_off) + (whole ? 0 : g_lo);
  const uint32_t* nos = (const uint32_t*)(h->slab + h->emitted.nos_off);
  uint32_t* io_sigs = (uint32_t*)slab_alloc(h, 4 * n_pairs + 4);
  uint32_t* io_nodes = (uint32_t*)slab_alloc(h, 4 * n_pairs + 4);
  uint32_t* es = es_ready ? es_ready : (uint32_t*)slab_alloc(h, 4 * ES_COUNT);
  uint4* d_new = outputs_on_device ? (uint4*)new_gates : (new_gates ? (uint4*)slab_alloc(h, 16 * G) : nullptr);
  uint32_t* d_order = outputs_on_device ? order_out : (order_out ? (uint32_t*)slab_alloc(h, 4 * G) : nullptr);
  uint32_t* d_wire = outputs_on_device && wire_of_node ? wire_of_node : (uint32_t*)slab_alloc(h, 4 * (size_t)node_bound);
  if (!d_wire || !io_sigs || !io_nodes || !es) return fail(h, C2A_ERR_NO_MEMORY, "scratch slab exhausted");
  if ((4 * n_pairs + 2048) > h->h_pinned_bytes) {
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    h->h_pinned_bytes = 4 * n_pairs + 8192;
    if (!cuda_ok(h, cudaHostAlloc((void**)&h->h_pinned, h->h_pinned_bytes, cudaHostAllocDefault), "cudaHostAlloc")) return C2A_ERR_CUDA;
  }
  const uint32_t* io_flag = nullptr;
  if (n_pairs) {
    uint32_t* stage = h->h_pinned + 256;
    if (h->emitted.nos_valid) {
      if (d_io_sigs_ready) io_sigs = const_cast<uint32_t*>(d_io_sigs_ready);
      else {
        if (n_in) memcpy(stage, input_signals, 4 * (size_t)n_in);
        if (n_out) memcpy(stage + n_in, output_signals, 4 * (size_t)n_out);
        cudaMemcpyAsync(io_sigs, stage, 4 * n_pairs, cudaMemcpyHostToDevice, s);
      }
      if (!es_ready) cudaMemsetAsync(es, 0, 4 * ES_COUNT, s);
      LAUNCH(h, k_ev_map_io, grid_for(h, (const void*)k_ev_map_io, kBlock, n_pairs), kBlock, io_sigs, (uint32_t)n_pairs, h->emitted.signal_bound, nos, io_nodes, es);
      io_flag = es + ES_IOBAD;  // read together with the build's final status: an unmapped signal yields node 0, which is harmless until then
    } else {  // sparse ids: map through the host emitter that produced the circuit
      if (n_in) c2a_signal_nodes(h->host_comp, input_signals, n_in, stage);
      if (n_out) c2a_signal_nodes(h->host_comp, output_signals, n_out, stage + n_in);
      for (size_t i = 0; i < n_pairs; ++i)
        if (stage[i] == 0) return fail(h, C2A_ERR_INVALID_ARGUMENT, "an input/output signal was never declared");
      cudaMemcpyAsync(io_nodes, stage, 4 * n_pairs, cudaMemcpyHostToDevice, s);
      if (!cuda_ok(h, cudaStreamSynchronize(s), "io upload")) return C2A_ERR_CUDA;
    }
  }
  h->emitted.wire = nullptr;
  bool identity = true;
  // host arrays: the copy-out is enqueued behind the pipeline and in FRONT of the build's status read, so the call synchronises once
  // (after an error the arrays hold unspecified contents - never anything behind their capacities)
  HostCopyOut copy_out;
  copy_out.new_gates_host = (uint4*)new_gates;
  copy_out.rest = [&]() {
    if (order_out && G) cudaMemcpyAsync(order_out, d_order, 4 * G, cudaMemcpyDeviceToHost, s);
    if (wire_of_node && node_bound) cudaMemcpyAsync(wire_of_node, d_wire, 4 * (size_t)node_bound, cudaMemcpyDeviceToHost, s);
  };
  st = build_core(h, p, d_gates, nullptr, nullptr, d_order, d_wire, d_new, wire_count, err_index, &identity, io_nodes, io_flag,
                  (whole && h->emitted.prod1_valid) ? (const uint32_t*)(h->slab + h->emitted.prod1_off) : nullptr, outputs_on_device ? nullptr : &copy_out);
  if (st == C2A_OK && whole) { h->emitted.wire = d_wire; h->emitted.identity = identity; }  // stay valid until the next call that carves the slab
  cudaStreamSynchronize(s);  // (build_core has synchronised already unless it left early)
  phases_collect(h);
  return st;
}

// selected signals -> wires (want_wires) or -> node ids, host lists in and out
static int emitted_signal_lookup(c2a_handle* h, const uint32_t* signals, uint64_t n, uint32_t* out_host, bool want_wires) {
  if (!h) return C2A_ERR_INVALID_ARGUMENT;
  if (!h->emitted.valid || (want_wires && !h->emitted.wire))
    return fail(h, C2A_ERR_INVALID_ARGUMENT, want_wires ? "no built circuit is resident on this handle" : "no emitted circuit is resident on this handle");
  if (n == 0) return C2A_OK;
  if (!signals || !out_host) return fail(h, C2A_ERR_INVALID_ARGUMENT, "null argument");
  if (!cuda_ok(h, cudaSetDevice(h->device), "cudaSetDevice")) return C2A_ERR_CUDA;
  cudaStream_t s = h->stream;
  const uint32_t* nos = h->emitted.nos_valid ? (const uint32_t*)(h->slab + h->emitted.nos_off) : nullptr;
  std::vector<uint32_t> nodes;
  const uint32_t* src = signals;
  if (!nos) {  // sparse ids: the host emitter that produced the circuit knows the nodes
    nodes.resize(n);
    c2a_signal_nodes(h->host_comp, signals, n, nodes.data());
    if (!want_wires) { memcpy(out_host, nodes.data(), 4 * n); return C2A_OK; }
    src = nodes.data();
  }
  // the event staging buffer is idle between calls: use it for the two lists (the slab holds the wire map itself)
  const size_t need = 2 * align256(4 * n);
  if (need > h->ev_bytes) {
    if (h->ev_buf) { cudaStreamSynchronize(s); cudaFree(h->ev_buf); h->ev_buf = nullptr; h->ev_bytes = 0; }
    if (!cuda_ok(h, cudaMalloc(&h->ev_buf, need), "cudaMalloc(signal list)")) { cudaGetLastError(); return C2A_ERR_NO_MEMORY; }
    h->ev_bytes = need;
  }
  uint32_t* d_sig = (uint32_t*)h->ev_buf;
  uint32_t* d_out = (uint32_t*)(h->ev_buf + align256(4 * n));
  if (!cuda_ok(h, cudaMemcpyAsync(d_sig, src, 4 * n, cudaMemcpyHostToDevice, s), "signal list H2D")) return C2A_ERR_CUDA;
  LAUNCH(h, k_sig_wires, grid_for(h, (const void*)k_sig_wires, kBlock, n), kBlock, d_sig, n, h->emitted.signal_bound, nos,
         want_wires ? h->emitted.wire : (const uint32_t*)nullptr, h->emitted.node_count + 1, d_out);
  if (!cuda_ok(h, cudaMemcpyAsync(out_host, d_out, 4 * n, cudaMemcpyDeviceToHost, s), "lookup D2H")) return C2A_ERR_CUDA;
  if (!cuda_ok(h, cudaStreamSynchronize(s), "signal lookup")) return C2A_ERR_CUDA;
  return cuda_ok(h, cudaGetLastError(), "k_sig_wires") ? C2A_OK : C2A_ERR_CUDA;
}
int c2a_emitted_signal_wires(c2a_handle* h, const uint32_t* signals, uint64_t n, uint32_t* wires_out) {
  return emitted_signal_lookup(h, signals, n, wires_out, true);
}
int c2a_emitted_signal_nodes(c2a_handle* h, const uint32_t* signals, uint64_t n, uint32_t* nodes_out) {
  return emitted_signal_lookup(h, signals, n, nodes_out, false);
}

int c2a_emitted_gather_device(c2a_handle* h, uint32_t* d_order, c2a_gate* d_new_gates, const uint64_t* d_counts, uint32_t rank, uint32_t world) {
  if (!h) return C2A_ERR_INVALID_ARGUMENT;
  if (!h->emitted.valid || !h->emitted.wire) return fail(h, C2A_ERR_INVALID_ARGUMENT, "no built circuit is resident on this handle");
  if (!d_new_gates || (d_counts && rank >= world)) return fail(h, C2A_ERR_INVALID_ARGUMENT, "bad argument");
  if (!h->emitted.identity && !d_order) return fail(h, C2A_ERR_INVALID_ARGUMENT, "the build's order is not the identity: pass its order array");
  if (!cuda_ok(h, cudaSetDevice(h->device), "cudaSetDevice")) return C2A_ERR_CUDA;
  const uint32_t G = (uint32_t)h->emitted.G;
  if (G) LAUNCH(h, k_gather_global, grid_for(h, (const void*)k_gather_global, kBlock, ((uint64_t)G + kGatherIlp - 1) / kGatherIlp), kBlock,
                (const uint4*)(h->slab + h->emitted.gates_off), d_order, h->emitted.identity ? 1u : 0u, G, h->emitted.wire, (uint4*)d_new_gates,
                (const unsigned long long*)d_counts, rank, world);
  return cuda_ok(h, cudaGetLastError(), "k_gather_global") ? C2A_OK : C2A_ERR_CUDA;
}

int c2a_emitted_signal_wires_device(c2a_handle* h, const uint32_t* d_signals, uint64_t n, uint32_t* d_wires_out) {
  if (!h) return C2A_ERR_INVALID_ARGUMENT;
  if (!h->emitted.valid || !h->emitted.wire) return fail(h, C2A_ERR_INVALID_ARGUMENT, "no built circuit is resident on this handle");
  if (!h->emitted.nos_valid) return fail(h, C2A_ERR_INVALID_ARGUMENT, "the signal -> node map is not resident (sparse signal ids): use c2a_emitted_signal_wires");
  if (n == 0) return C2A_OK;
  if (!d_signals || !d_wires_out) return fail(h, C2A_ERR_INVALID_ARGUMENT, "null argument");
  if (!cuda_ok(h, cudaSetDevice(h->device), "cudaSetDevice")) return C2A_ERR_CUDA;
  LAUNCH(h, k_sig_wires, grid_for(h, (const void*)k_sig_wires, kBlock, n), kBlock, d_signals, n, h->emitted.signal_bound, (const uint32_t*)(h->slab + h->emitted.nos_off),
         h->emitted.wire, h->emitted.node_count + 1, d_wires_out);
  return cuda_ok(h, cudaGetLastError(), "k_sig_wires") ? C2A_OK : C2A_ERR_CUDA;
}

int c2a_emitted_build_circuit(c2a_handle* h, const uint32_t* input_signals, uint32_t n_in, const uint32_t* output_signals, uint32_t n_out,
                              uint32_t* order_out, uint32_t* wire_of_node, c2a_gate* new_gates, uint32_t* wire_count, uint64_t* err_index) {
  return emitted_build_impl(h, input_signals, n_in, output_signals, n_out, order_out, wire_of_node, new_gates, wire_count, err_index, false);
}
int c2a_emitted_build_circuit_device(c2a_handle* h, const uint32_t* input_signals, uint32_t n_in, const uint32_t* output_signals, uint32_t n_out,
                                     uint32_t* d_order_out, uint32_t* d_wire_of_node, c2a_gate* d_new_gates, uint32_t* wire_count,
                                     uint64_t* err_index) {
  return emitted_build_impl(h, input_signals, n_in, output_signals, n_out, d_order_out, d_wire_of_node, d_new_gates, wire_count, err_index, true);
}
int c2a_emitted_build_range_device(c2a_handle* h, uint64_t gate_lo, uint64_t gate_hi, const uint32_t* input_signals, uint32_t n_in,
                                   const uint32_t* output_signals, uint32_t n_out, uint32_t* d_order_out, uint32_t* d_wire_of_node, c2a_gate* d_new_gates,
                                   uint32_t* wire_count, uint64_t* err_index) {
  return emitted_build_impl(h, input_signals, n_in, output_signals, n_out, d_order_out, d_wire_of_node, d_new_gates, wire_count, err_index, true, gate_lo, gate_hi);
}
const c2a_gate* c2a_emitted_gates_device(c2a_handle* h) {
  return (h && h->emitted.valid) ? (const c2a_gate*)(h->slab + h->emitted.gates_off) : nullptr;
}

}  // extern "C"
