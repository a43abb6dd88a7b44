// c2a_eval.cuh — level-parallel u32 circuit evaluator.  Included by c2a_device.cu.
//
// Restates the test-side simulator of the reference (tests/integration.rs:90-119, `ArithmeticGate::execute` on
// u32, driven by sim-circuit's straight-line executor): gates are executed in the order given, each reads two wires
// that must already hold a value and writes one.  Native-op semantics of a release build: add / sub / mul / pow wrap,
// shifts by >= 32 give 0, division or remainder by zero stops the run (Rust panics).
//
// Parallel form: the circuit must be single-assignment (a wire is written by at most one gate and never an initially
// set wire - what build_circuit produces); then the value of every wire is order-independent and the gates can be
// evaluated level by level over the Kahn levels (K3/K4).  The sequential executor's failure point is reproduced
// exactly: *err_index = the smallest gate position that reads an unset / not-yet-written wire or divides by zero
// (a failing gate writes 0, so garbage can only flow to larger positions, which cannot lower the minimum).
#pragma once

namespace c2a {

enum { EV_ERR = 0, EV_FLAGS = 1, EV_COUNT = 4 };
enum { EVF_BAD = 1, EVF_MULTI = 2 };

__device__ __forceinline__ bool eval_op_u32(uint32_t op, uint32_t x, uint32_t y, uint32_t* r) {
  switch (op) {
    case C2A_AAdd: *r = x + y; return true;
    case C2A_ADiv: case C2A_AIntDiv: if (!y) return false; *r = x / y; return true;
    case C2A_AEq: *r = x == y; return true;
    case C2A_AGEq: *r = x >= y; return true;
    case C2A_AGt: *r = x > y; return true;
    case C2A_ALEq: *r = x <= y; return true;
    case C2A_ALt: *r = x < y; return true;
    case C2A_AMul: *r = x * y; return true;
    case C2A_ANeq: *r = x != y; return true;
    case C2A_ASub: *r = x - y; return true;
    case C2A_AXor: *r = x ^ y; return true;
    case C2A_APow: { uint32_t acc = 1, base = x, e = y; while (e) { if (e & 1) acc *= base; e >>= 1; base *= base; } *r = acc; return true; }
    case C2A_AMod: if (!y) return false; *r = x % y; return true;
    case C2A_AShiftL: *r = y < 32 ? x << y : 0u; return true;
    case C2A_AShiftR: *r = y < 32 ? x >> y : 0u; return true;
    case C2A_ABoolOr: *r = (x != 0 || y != 0); return true;
    case C2A_ABoolAnd: *r = (x != 0 && y != 0); return true;
    case C2A_ABitOr: *r = x | y; return true;
    case C2A_ABitAnd: *r = x & y; return true;
  }
  return false;
}

// writer[w] = the gate that writes wire w (single assignment is checked here)
__global__ void __launch_bounds__(kBlock) k_eval_writers(const uint4* __restrict__ gates, uint32_t G, uint32_t W, const uint8_t* __restrict__ has,
                                                         uint32_t* __restrict__ writer, uint32_t* __restrict__ sc) {
  uint32_t f = 0;
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < G; g += gridDim.x * kBlock) {
    uint32_t o = ldg_stream(gates + g).w;
    if (o >= W) { f |= EVF_BAD; continue; }
    if (has[o]) f |= EVF_MULTI;  // an initially set wire is written again
    if (atomicCAS(writer + o, kNone, g) != kNone) f |= EVF_MULTI;
  }
  f = warp_or(f);
  if ((threadIdx.x & 31) == 0 && f) atomicOr(sc + EV_FLAGS, f);
}

// Gates that the straight-line executor cannot run (an operand is unset, out of range, or written at a later / the same
// position) are recorded and their operands redirected to the always-set dummy wire W, which keeps the level graph acyclic.
__global__ void __launch_bounds__(kBlock) k_eval_check(uint4* __restrict__ gates, uint32_t G, uint32_t W, const uint8_t* __restrict__ has,
                                                       const uint32_t* __restrict__ writer, uint8_t* __restrict__ failed, uint32_t* __restrict__ sc) {
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < G; g += gridDim.x * kBlock) {
    uint4 gt = gates[g];
    bool ok = gt.x < C2A_GATE_TYPE_COUNT;
    uint32_t in[2] = {gt.y, gt.z};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      uint32_t w = in[k];
      if (w >= W) { ok = false; continue; }
      uint32_t p = writer[w];
      if (p == kNone ? !has[w] : p >= g) ok = false;
    }
    failed[g] = ok ? 0 : 1;
    if (!ok) {
      atomicMin(sc + EV_ERR, g);
      gates[g] = make_uint4(gt.x, W, W, gt.w);
    }
  }
}

__device__ __forceinline__ void eval_one(const uint4* __restrict__ gates, uint32_t g, const uint8_t* __restrict__ failed, uint32_t* __restrict__ val,
                                         uint8_t* __restrict__ has, uint32_t* __restrict__ sc) {
  uint4 gt = gates[g];
  uint32_t r = 0;
  // values written in an earlier level: plain loads would be fine across launches; inside the single-CTA multi-level
  // kernel they must not be served from a stale L1 line -> ld.cg
  bool ok = !failed[g] && eval_op_u32(gt.x, __ldcg(val + gt.y), __ldcg(val + gt.z), &r);
  if (!ok) { atomicMin(sc + EV_ERR, g); r = 0; }
  __stcg(val + gt.w, r);
  has[gt.w] = 1;
}

__global__ void __launch_bounds__(kBlock) k_eval_level(const uint4* __restrict__ gates, const uint32_t* __restrict__ level_order, uint32_t lo, uint32_t hi,
                                                       const uint8_t* __restrict__ failed, uint32_t* __restrict__ val, uint8_t* __restrict__ has,
                                                       uint32_t* __restrict__ sc) {
  for (uint32_t i = lo + blockIdx.x * kBlock + threadIdx.x; i < hi; i += gridDim.x * kBlock) eval_one(gates, level_order[i], failed, val, has, sc);
}

// consecutive narrow levels [l0, l1): one CTA walks them with a block barrier in between (no launch per level)
__global__ void __launch_bounds__(1024) k_eval_levels_narrow(const uint4* __restrict__ gates, const uint32_t* __restrict__ level_order,
                                                             const uint32_t* __restrict__ level_off, uint32_t l0, uint32_t l1,
                                                             const uint8_t* __restrict__ failed, uint32_t* __restrict__ val, uint8_t* __restrict__ has,
                                                             uint32_t* __restrict__ sc) {
  for (uint32_t l = l0; l < l1; ++l) {
    uint32_t lo = level_off[l], hi = level_off[l + 1];
    for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) eval_one(gates, level_order[i], failed, val, has, sc);
    __syncthreads();  // orders this CTA's global writes before the next level's reads (same CTA, L2-coherent ld.cg / st.cg)
  }
}

}  // namespace c2a

using namespace c2a;

extern "C" int c2a_evaluate(c2a_handle* h, const c2a_gate* gates, uint64_t G, uint32_t wire_count, uint32_t* values, uint8_t* has, uint64_t* err_index) {
  int st = check_sizes(h, G, wire_count);
  if (st) return st;
  if ((G && !gates) || (wire_count && (!values || !has))) return fail(h, C2A_ERR_INVALID_ARGUMENT, "null argument");
  if (wire_count >= kOutPending - 1) return fail(h, C2A_ERR_INVALID_ARGUMENT, "wire_count too large");
  phases_clear(h);
  slab_reset(h);
  const uint32_t W = wire_count, NB = W + 1;  // + the dummy wire
  size_t need = kahn_scratch_bytes(G, NB) + align256(16 * G) + align256(4 * G) + align256(4 * (G + 2)) + align256(4 * (size_t)NB) + align256(NB) +
                align256(4 * (size_t)NB) + align256(G) + align256(4 * EV_COUNT);
  if (!slab_reserve(h, need)) return C2A_ERR_NO_MEMORY;
  KahnBuffers b;
  bool ok = kahn_carve(h, G, NB, &b);
  uint4* d_gates = (uint4*)slab_alloc(h, 16 * G);
  uint32_t* d_lo = (uint32_t*)slab_alloc(h, 4 * G);
  uint32_t* d_off = (uint32_t*)slab_alloc(h, 4 * (G + 2));
  uint32_t* d_val = (uint32_t*)slab_alloc(h, 4 * (size_t)NB);
  uint8_t* d_has = (uint8_t*)slab_alloc(h, NB);
  uint32_t* d_writer = (uint32_t*)slab_alloc(h, 4 * (size_t)NB);
  uint8_t* d_failed = (uint8_t*)slab_alloc(h, G);
  uint32_t* sc = (uint32_t*)slab_alloc(h, 4 * EV_COUNT);
  if (!ok || !d_gates || !d_lo || !d_off || !d_val || !d_has || !d_writer || !d_failed || !sc) return fail(h, C2A_ERR_NO_MEMORY, "scratch slab exhausted");
  cudaStream_t s = h->stream;
  uint32_t* hp = h->h_pinned;
  phase_begin(h, "h2d");
  if (G) cudaMemcpyAsync(d_gates, gates, 16 * G, cudaMemcpyHostToDevice, s);
  if (W) { cudaMemcpyAsync(d_val, values, 4 * (size_t)W, cudaMemcpyHostToDevice, s); cudaMemcpyAsync(d_has, has, W, cudaMemcpyHostToDevice, s); }
  phase_end(h);
  cudaMemsetAsync(d_val + W, 0, 4, s);
  cudaMemsetAsync(d_has + W, 1, 1, s);
  cudaMemsetAsync(d_writer, 0xFF, 4 * (size_t)NB, s);
  cudaMemsetAsync(d_off, 0, 4 * (G + 2), s);
  hp[200] = kNone; hp[201] = 0; hp[202] = 0; hp[203] = 0;
  cudaMemcpyAsync(sc, hp + 200, 4 * EV_COUNT, cudaMemcpyHostToDevice, s);
  if (G) {
    phase_begin(h, "k_eval_writers");
    LAUNCH(h, k_eval_writers, grid_for(h, (const void*)k_eval_writers, kBlock, G), kBlock, d_gates, (uint32_t)G, W, d_has, d_writer, sc);
    phase_end(h);
    phase_begin(h, "k_eval_check");
    LAUNCH(h, k_eval_check, grid_for(h, (const void*)k_eval_check, kBlock, G), kBlock, d_gates, (uint32_t)G, W, d_has, d_writer, d_failed, sc);
    phase_end(h);
  }
  cudaMemcpyAsync(hp + 200, sc, 4 * EV_COUNT, cudaMemcpyDeviceToHost, s);
  if (!cuda_ok(h, cudaStreamSynchronize(s), "evaluate setup")) return C2A_ERR_CUDA;
  if (hp[200 + EV_FLAGS] & EVF_BAD) return fail(h, C2A_ERR_INVALID_ARGUMENT, "a gate writes a wire >= wire_count (%u)", W);
  if (hp[200 + EV_FLAGS] & EVF_MULTI) return fail(h, C2A_ERR_INVALID_ARGUMENT, "the circuit is not single-assignment (a wire is written twice, or an initially set wire is written)");
  uint32_t nl = 0;
  uint64_t cyc = 0;
  st = kahn_core(h, d_gates, (uint32_t)G, NB, b, d_lo, d_off, (uint32_t)std::min<uint64_t>(G, 0xFFFFFFFEull), &nl, &cyc);
  if (st != C2A_OK) { phases_collect(h); return st == C2A_ERR_CYCLIC_DEPENDENCY ? fail(h, C2A_ERR_CUDA, "internal: level graph has a cycle") : st; }
  std::vector<uint32_t> off((size_t)nl + 1, 0);
  if (nl) cudaMemcpyAsync(off.data(), d_off, 4 * ((size_t)nl + 1), cudaMemcpyDeviceToHost, s);
  if (!cuda_ok(h, cudaStreamSynchronize(s), "levels D2H")) return C2A_ERR_CUDA;
  if (nl) off[nl] = (uint32_t)G;
  // d_off[nl] on the device must also close the last level for the narrow kernel
  if (nl) cudaMemcpyAsync(d_off + nl, off.data() + nl, 4, cudaMemcpyHostToDevice, s);
  phase_begin(h, "k_eval_level");
  const uint32_t kNarrow = 2048;
  for (uint32_t l = 0; l < nl;) {
    uint32_t n = off[l + 1] - off[l];
    if (n > kNarrow) {
      LAUNCH(h, k_eval_level, grid_for(h, (const void*)k_eval_level, kBlock, n), kBlock, d_gates, d_lo, off[l], off[l + 1], d_failed, d_val, d_has, sc);
      ++l;
    } else {
      uint32_t l1 = l + 1;
      while (l1 < nl && off[l1 + 1] - off[l1] <= kNarrow) ++l1;
      LAUNCH(h, k_eval_levels_narrow, 1, 1024, d_gates, d_lo, d_off, l, l1, d_failed, d_val, d_has, sc);
      l = l1;
    }
  }
  phase_end(h);
  cudaMemcpyAsync(hp + 200, sc, 4 * EV_COUNT, cudaMemcpyDeviceToHost, s);
  phase_begin(h, "d2h");
  if (W) { cudaMemcpyAsync(values, d_val, 4 * (size_t)W, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync(has, d_has, W, cudaMemcpyDeviceToHost, s); }
  phase_end(h);
  if (!cuda_ok(h, cudaStreamSynchronize(s), "evaluate D2H")) return C2A_ERR_CUDA;
  if (!cuda_ok(h, cudaGetLastError(), "evaluate kernels")) return C2A_ERR_CUDA;
  phases_collect(h);
  if (hp[200 + EV_ERR] != kNone) {
    if (err_index) *err_index = hp[200 + EV_ERR];
    return fail(h, C2A_ERR_EVALUATION, "gate %u reads a wire without a value or divides by zero", hp[200 + EV_ERR]);
  }
  return C2A_OK;
}
