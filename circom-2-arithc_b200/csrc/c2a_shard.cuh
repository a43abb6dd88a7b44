// c2a_shard.cuh — where does ONE gate vector split into independent component subtrees?  Device version of
// sharding.find_cuts / plan_shards (which walk the whole gate vector with numpy on the host: seconds at 10 M gates).
//
// A cut at gate index c (0 < c < G) is valid iff no dependency edge (src/compiler.rs:408-421) and no non-I/O node crosses it: then
// the reference's DFS started from a root < c never reaches a gate >= c, first-seen wire numbering of the right part continues where
// the left part stopped, and the per-shard results concatenate to the single-GPU result bit for bit (circom-2-arithc_b200/sharding.py).
// Every crossing object is an interval (lo, hi] of forbidden cuts.  With H[i] = the largest hi among the intervals that start at lo = i
// (H[i] = i when there is none), c is valid iff max(H[0..c-1]) == c - 1: ONE running maximum.
//   k_cut_span      first / last gate using every node (RED.MIN / RED.MAX)
//   k_cut_io        I/O nodes are numbered from the shared lists, not first-seen: their spans do not count
//   k_cut_h_nodes   H[first[X]] = max(.., last[X]) for every other node X
//   k_cut_h_deps    H[min(g, d)] = max(.., max(g, d)) for every dependency edge g -> d (d = last producer of an operand)
//   k_cut_tilemax / k_cut_carry / k_cut_pick   running maximum in 4096-gate tiles; the valid cuts next to the targets k*G/world
#pragma once

namespace c2a {

constexpr int kCutItems = 16, kCutTile = kBlock * kCutItems;

__global__ void __launch_bounds__(kBlock) k_cut_span(const uint4* __restrict__ gates, uint32_t G, uint32_t node_bound, uint32_t* __restrict__ first,
                                                     uint32_t* __restrict__ last1, uint32_t* __restrict__ flags) {
  bool bad = false;
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < G; g += gridDim.x * kBlock) {
    const uint4 gt = ldg_stream(gates + g);
    const uint32_t nd[3] = {gt.y, gt.z, gt.w};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (nd[j] >= node_bound) { bad = true; continue; }
      if (first[nd[j]] > g) atomicMin(first + nd[j], g);
      if (last1[nd[j]] < g + 1) atomicMax(last1 + nd[j], g + 1);
    }
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flags, 1u);
}
__global__ void __launch_bounds__(kBlock) k_cut_io(const uint32_t* __restrict__ io, uint32_t n, uint32_t node_bound, uint32_t* __restrict__ last1) {
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock)
    if (io[i] < node_bound) last1[io[i]] = 0;
}
__global__ void __launch_bounds__(kBlock) k_cut_h_nodes(const uint32_t* __restrict__ first, const uint32_t* __restrict__ last1, uint32_t node_bound,
                                                        uint32_t* __restrict__ H) {
  for (uint32_t x = blockIdx.x * kBlock + threadIdx.x; x < node_bound; x += gridDim.x * kBlock) {
    const uint32_t l1 = last1[x];
    if (l1 && l1 - 1 > first[x]) atomicMax(H + first[x], l1 - 1);
  }
}
__global__ void __launch_bounds__(kBlock) k_cut_h_deps(const uint4* __restrict__ gates, uint32_t G, uint32_t node_bound, const uint32_t* __restrict__ prod1,
                                                       uint32_t* __restrict__ H) {
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < G; g += gridDim.x * kBlock) {
    const uint4 gt = ldg_stream(gates + g);
    const uint32_t nd[2] = {gt.y, gt.z};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (nd[j] >= node_bound) continue;
      const uint32_t p1 = __ldg(prod1 + nd[j]);
      if (!p1 || p1 - 1 == g) continue;
      const uint32_t d = p1 - 1, lo = min(g, d), hi = max(g, d);
      if (H[lo] < hi) atomicMax(H + lo, hi);
    }
  }
}
// tile maxima of H
__global__ void __launch_bounds__(kBlock) k_cut_tilemax(const uint32_t* __restrict__ H, uint32_t G, uint32_t* __restrict__ tmax) {
  __shared__ uint32_t s_m[kBlock / 32];
  const uint32_t base = blockIdx.x * kCutTile;
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < kCutItems; ++i) {
    const uint32_t g = base + i * kBlock + threadIdx.x;
    if (g < G) m = max(m, H[g]);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kBlock / 32; ++w) t = max(t, s_m[w]);
    tmax[blockIdx.x] = t;
  }
}
// exclusive running maximum over the tile maxima (one CTA; <= 2^17 tiles at the 2^29-gate limit)
__global__ void __launch_bounds__(kBlock) k_cut_carry(uint32_t* __restrict__ tmax, uint32_t tiles) {
  __shared__ uint32_t s_part[kBlock];
  const uint32_t per = (tiles + kBlock - 1) / kBlock, lo = min(tiles, threadIdx.x * per), hi = min(tiles, lo + per);
  uint32_t m = 0;
  for (uint32_t t = lo; t < hi; ++t) m = max(m, tmax[t]);
  s_part[threadIdx.x] = m;
  __syncthreads();
  uint32_t carry = 0;
  for (uint32_t j = 0; j < threadIdx.x; ++j) carry = max(carry, s_part[j]);
  for (uint32_t t = lo; t < hi; ++t) { const uint32_t v = tmax[t]; tmax[t] = carry; carry = max(carry, v); }
}
// valid cuts: c in [1, G) with max(H[0..c-1]) == c - 1.  res[0] = their number; for target k (1..world-1) at t_k = k*G/world:
// res[2k] = the largest valid cut < t_k (0 = none), res[2k+1] = the smallest valid cut >= t_k (~0 = none).
__global__ void __launch_bounds__(kBlock) k_cut_pick(const uint32_t* __restrict__ H, uint32_t G, const uint32_t* __restrict__ carry, uint32_t world,
                                                     uint32_t* res) {
  __shared__ uint32_t s_scan[kBlock];
  const uint32_t base = blockIdx.x * kCutTile + threadIdx.x * kCutItems;  // kCutItems consecutive gates per thread
  uint32_t v[kCutItems], m = 0;
#pragma unroll
  for (int i = 0; i < kCutItems; ++i) { v[i] = base + i < G ? H[base + i] : 0u; m = max(m, v[i]); }
  s_scan[threadIdx.x] = m;
  __syncthreads();
  uint32_t run = carry[blockIdx.x];
  for (uint32_t j = 0; j < threadIdx.x; ++j) run = max(run, s_scan[j]);  // (256 shared loads per thread: this kernel runs once per plan)
  uint32_t nvalid = 0;
#pragma unroll
  for (int i = 0; i < kCutItems; ++i) {
    const uint32_t g = base + i;
    run = max(run, v[i]);
    const uint32_t c = g + 1;  // the cut behind gate g
    if (g < G && c < G && run == g) {
      ++nvalid;
      // every target may have c as its nearest cut on one side (words only move one way: read before the atomic)
      for (uint32_t k = 1; k < world; ++k) {
        const uint32_t t = (uint32_t)((unsigned long long)k * G / world);
        if (c < t) { if (res[2 * k] < c) atomicMax(res + 2 * k, c); }
        else if (res[2 * k + 1] > c) atomicMin(res + 2 * k + 1, c);
      }
    }
  }
  nvalid = warp_sum(nvalid);
  if ((threadIdx.x & 31) == 0 && nvalid) atomicAdd(res, nvalid);
}

}  // namespace c2a

using namespace c2a;

extern "C" {

int c2a_plan_shards_device(c2a_handle* h, const c2a_gate* d_gates, uint64_t G, uint32_t node_bound, const uint32_t* input_nodes, uint32_t n_in,
                           const uint32_t* output_nodes, uint32_t n_out, uint32_t world, uint64_t* bounds_out, uint32_t* n_shards) {
  int st = check_sizes(h, G, node_bound);
  if (st) return st;
  if (!bounds_out || !n_shards || world == 0 || world > 1024 || (G && !d_gates)) return fail(h, C2A_ERR_INVALID_ARGUMENT, "bad argument");
  *n_shards = 1;
  bounds_out[0] = 0;
  bounds_out[1] = G;
  if (world == 1 || G < 2) { if (world > 1) return C2A_OK; *n_shards = 1; return C2A_OK; }
  cudaStream_t s = h->stream;
  const uint32_t Gn = (uint32_t)G, tiles = (Gn + kCutTile - 1) / kCutTile;
  const size_t n_io = (size_t)n_in + n_out;
  // own grow-only allocation: the gate vector may live in the handle's slab (an emitted circuit), which must not move
  const size_t bytes = 3 * align256(4 * (size_t)node_bound) + align256(4 * G) + align256(4 * (size_t)tiles + 4) + align256(4 * n_io + 4) + align256(8 * (size_t)world + 64);
  if (bytes > h->plan_bytes) {
    if (h->plan_buf) { cudaStreamSynchronize(s); cudaFree(h->plan_buf); h->plan_buf = nullptr; h->plan_bytes = 0; }
    if (!cuda_ok(h, cudaMalloc(&h->plan_buf, bytes + bytes / 8), "cudaMalloc(shard plan)")) { cudaGetLastError(); return C2A_ERR_NO_MEMORY; }
    h->plan_bytes = bytes + bytes / 8;
  }
  char* q = h->plan_buf;
  auto take = [&](size_t b) { char* r = q; q += align256(b); return r; };
  uint32_t* first = (uint32_t*)take(4 * (size_t)node_bound);
  uint32_t* last1 = (uint32_t*)take(4 * (size_t)node_bound);
  uint32_t* prod1 = (uint32_t*)take(4 * (size_t)node_bound);
  uint32_t* H = (uint32_t*)take(4 * G);
  uint32_t* tmax = (uint32_t*)take(4 * (size_t)tiles + 4);
  uint32_t* d_io = (uint32_t*)take(4 * n_io + 4);
  uint32_t* res = (uint32_t*)take(8 * (size_t)world + 64);  // [0] count, [1] flags, [2k], [2k+1]
  std::vector<uint32_t> io(n_io), init(2 * (size_t)world + 2, 0u);
  if (n_in) memcpy(io.data(), input_nodes, 4 * (size_t)n_in);
  if (n_out) memcpy(io.data() + n_in, output_nodes, 4 * (size_t)n_out);
  for (uint32_t k = 1; k < world; ++k) init[2 * k + 1] = 0xFFFFFFFFu;
  phases_clear(h);
  phase_begin(h, "k_cut_plan");
  if (n_io) cudaMemcpyAsync(d_io, io.data(), 4 * n_io, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(res, init.data(), 4 * init.size(), cudaMemcpyHostToDevice, s);
  cudaMemsetAsync(first, 0xFF, 4 * (size_t)node_bound, s);
  cudaMemsetAsync(last1, 0, 4 * (size_t)node_bound, s);
  cudaMemsetAsync(prod1, 0, 4 * (size_t)node_bound, s);
  LAUNCH(h, k_iota, grid_for(h, (const void*)k_iota, kBlock, G), kBlock, H, Gn);
  LAUNCH(h, k_cut_span, grid_for(h, (const void*)k_cut_span, kBlock, G), kBlock, (const uint4*)d_gates, Gn, node_bound, first, last1, res + 1);
  LAUNCH(h, k_producer, grid_for(h, (const void*)k_producer, kBlock, G), kBlock, (const uint4*)d_gates, Gn, node_bound, prod1, tmax /* its F_BAD word: unused here */);
  if (n_io) LAUNCH(h, k_cut_io, grid_for(h, (const void*)k_cut_io, kBlock, n_io), kBlock, d_io, (uint32_t)n_io, node_bound, last1);
  LAUNCH(h, k_cut_h_nodes, grid_for(h, (const void*)k_cut_h_nodes, kBlock, node_bound), kBlock, first, last1, node_bound, H);
  LAUNCH(h, k_cut_h_deps, grid_for(h, (const void*)k_cut_h_deps, kBlock, G), kBlock, (const uint4*)d_gates, Gn, node_bound, prod1, H);
  LAUNCH(h, k_cut_tilemax, tiles, kBlock, H, Gn, tmax);
  LAUNCH(h, k_cut_carry, 1, kBlock, tmax, tiles);
  LAUNCH(h, k_cut_pick, tiles, kBlock, H, Gn, tmax, world, res);
  phase_end(h);
  std::vector<uint32_t> out(2 * (size_t)world + 2);
  cudaMemcpyAsync(out.data(), res, 4 * out.size(), cudaMemcpyDeviceToHost, s);
  if (!cuda_ok(h, cudaStreamSynchronize(s), "shard plan sync")) return C2A_ERR_CUDA;
  if (!cuda_ok(h, cudaGetLastError(), "shard plan kernels")) return C2A_ERR_CUDA;
  phases_collect(h);
  if (out[1]) return fail(h, C2A_ERR_INVALID_ARGUMENT, "a gate references a node id >= node_bound (%u)", node_bound);
  if (out[0] < world - 1) return C2A_OK;  // fewer independent subtrees than ranks: replicas only (*n_shards stays 1)
  // the rule of sharding.plan_shards: per target the nearer of the two neighbouring cuts (the lower one on a tie) that lies behind
  // the previous bound
  std::vector<uint64_t> b(1, 0);
  for (uint32_t k = 1; k < world; ++k) {
    const uint64_t target = (uint64_t)k * G / world;
    uint64_t best = 0;
    bool have = false;
    for (int side = 0; side < 2; ++side) {
      const uint32_t c = out[2 * k + side];
      if ((side == 0 && c == 0) || (side == 1 && c == 0xFFFFFFFFu) || c <= b.back()) continue;
      const uint64_t dist = c > target ? c - target : target - c, bdist = best > target ? best - target : target - best;
      if (!have || dist < bdist) { best = c; have = true; }
    }
    if (!have) return C2A_OK;  // (sharding.plan_shards would look further right; with balanced targets this means: does not split evenly)
    b.push_back(best);
  }
  b.push_back(G);
  for (size_t i = 0; i + 1 < b.size(); ++i) if (b[i + 1] <= b[i]) return C2A_OK;
  for (size_t i = 0; i < b.size(); ++i) bounds_out[i] = b[i];
  *n_shards = world;
  return C2A_OK;
}

}  // extern "C"
