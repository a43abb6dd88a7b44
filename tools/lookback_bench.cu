// lookback_bench.cu — micro-benchmark of the single-pass (decoupled look-back) exclusive scan variants used by
// k_scan_u32 / k_wire_scan.  Developer tool, not part of the product:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o lookback_bench tools/lookback_bench.cu && ./lookback_bench
// Variants: tile index from a global ticket vs blockIdx.x; look-back window of 32 * W tiles; ITEMS per thread.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int kBlock = 256;
__device__ __forceinline__ unsigned long long ld_st(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_st(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
constexpr unsigned long long kAgg = 1ull << 32, kInc = 2ull << 32;

template <int W>
__device__ __forceinline__ uint32_t lookback(const unsigned long long* st, uint32_t tile, int lane) {
  uint32_t prefix = 0;
  long long p = (long long)tile - 1;
  while (true) {
    unsigned long long v[W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
      long long idx = p - lane - 32 * j;
      v[j] = idx >= 0 ? ld_st(st + idx) : kInc;
    }
#pragma unroll
    for (int j = 0; j < W; ++j) {
      long long idx = p - lane - 32 * j;
      while (true) {
        uint32_t s = (uint32_t)(v[j] >> 32);
        uint32_t empty = __ballot_sync(0xFFFFFFFFu, s == 0), inc = __ballot_sync(0xFFFFFFFFu, s == 2);
        uint32_t val = (uint32_t)v[j];
        if (inc) {
          int first = __ffs(inc) - 1;
          if (!(empty & ((1u << first) - 1u))) return prefix + warp_sum(lane <= first ? val : 0u);
        } else if (!empty) {
          prefix += warp_sum(val);
          break;
        }
        if (s == 0) v[j] = ld_st(st + idx);
      }
    }
    p -= 32 * W;
  }
}

// MODE 0: ticket, 1: blockIdx.  NOLB: skip the look-back (wrong result; isolates its cost)
template <int ITEMS, int W, int MODE, bool NOLB>
__global__ void __launch_bounds__(kBlock) k_scan(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint32_t n, unsigned long long* __restrict__ st,
                                                 uint32_t* __restrict__ ticket) {
  __shared__ uint32_t s_mem[12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t tile;
  if (MODE == 0) {
    if (threadIdx.x == 0) s_mem[9] = atomicAdd(ticket, 1u);
    __syncthreads();
    tile = s_mem[9];
  } else tile = blockIdx.x;
  uint32_t base = tile * (kBlock * ITEMS) + threadIdx.x * ITEMS;
  uint32_t v[ITEMS];
  if (base + ITEMS <= n) {
#pragma unroll
    for (int i = 0; i < ITEMS; i += 4) {
      uint4 x = *reinterpret_cast<const uint4*>(src + base + i);
      v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) v[i] = base + i < n ? src[base + i] : 0u;
  }
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) sum += v[i];
  uint32_t incl = warp_incl_scan(sum, lane);
  if (lane == 31) s_mem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < 8 ? s_mem[lane] : 0u;
    uint32_t wi = warp_incl_scan(w, lane);
    uint32_t agg = __shfl_sync(0xFFFFFFFFu, wi, 7);
    if (lane < 8) s_mem[lane] = wi - w;
    uint32_t prefix = 0;
    if (!NOLB) {
      if (tile == 0) {
        if (lane == 0) st_st(st, kInc | agg);
      } else {
        if (lane == 0) st_st(st + tile, kAgg | agg);
        prefix = lookback<W>(st, tile, lane);
        if (lane == 0) st_st(st + tile, kInc | (unsigned long long)(prefix + agg));
      }
    }
    if (lane == 0) s_mem[8] = prefix;
  }
  __syncthreads();
  uint32_t ex = s_mem[8] + s_mem[warp] + (incl - sum);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) { uint32_t t = v[i]; v[i] = ex; ex += t; }
  if (base + ITEMS <= n) {
#pragma unroll
    for (int i = 0; i < ITEMS; i += 4) *reinterpret_cast<uint4*>(dst + base + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) if (base + i < n) dst[base + i] = v[i];
  }
}

template <int ITEMS, int W, int MODE, bool NOLB>
static void run(const char* name, const uint32_t* src, uint32_t* dst, uint32_t n, unsigned long long* st, uint32_t* ticket, const std::vector<uint32_t>& ref) {
  uint32_t tiles = (n + kBlock * ITEMS - 1) / (kBlock * ITEMS);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e9f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaMemset(st, 0, 8 * (size_t)tiles);
    cudaMemset(ticket, 0, 4);
    cudaEventRecord(a);
    k_scan<ITEMS, W, MODE, NOLB><<<tiles, kBlock>>>(src, dst, n, st, ticket);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (rep) best = ms < best ? ms : best;
  }
  bool ok = true;
  if (!NOLB) {
    std::vector<uint32_t> out(n);
    cudaMemcpy(out.data(), dst, 4 * (size_t)n, cudaMemcpyDeviceToHost);
    for (uint32_t i = 0; i < n; i += 9973) if (out[i] != ref[i]) { ok = false; break; }
  }
  printf("%-34s tiles %6u  %8.1f us  %7.1f GB/s  %s\n", name, tiles, best * 1e3, 8.0 * n / (best * 1e-3) / 1e9, NOLB ? "(no look-back)" : ok ? "ok" : "WRONG");
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("  cuda error: %s\n", cudaGetErrorString(e));
}

int main() {
  const uint32_t n = 10018305;
  std::vector<uint32_t> hsrc(n), ref(n);
  uint32_t acc = 0;
  for (uint32_t i = 0; i < n; ++i) { hsrc[i] = (i * 2654435761u >> 29) & 3; ref[i] = acc; acc += hsrc[i]; }
  uint32_t *src, *dst, *ticket;
  unsigned long long* st;
  cudaMalloc(&src, 4 * (size_t)n);
  cudaMalloc(&dst, 4 * (size_t)n + 4);
  cudaMalloc(&st, 8 * (size_t)(n / 256 + 2));
  cudaMalloc(&ticket, 4);
  cudaMemcpy(src, hsrc.data(), 4 * (size_t)n, cudaMemcpyHostToDevice);
  run<8, 1, 0, false>("ticket  items 8  win 32", src, dst, n, st, ticket, ref);
  run<8, 1, 1, false>("blockIdx items 8  win 32", src, dst, n, st, ticket, ref);
  run<8, 4, 1, false>("blockIdx items 8  win 128", src, dst, n, st, ticket, ref);
  run<16, 1, 0, false>("ticket  items 16 win 32", src, dst, n, st, ticket, ref);
  run<16, 1, 1, false>("blockIdx items 16 win 32", src, dst, n, st, ticket, ref);
  run<16, 4, 1, false>("blockIdx items 16 win 128", src, dst, n, st, ticket, ref);
  run<16, 4, 0, false>("ticket  items 16 win 128", src, dst, n, st, ticket, ref);
  run<4, 1, 1, false>("blockIdx items 4  win 32", src, dst, n, st, ticket, ref);
  run<4, 4, 1, false>("blockIdx items 4  win 128", src, dst, n, st, ticket, ref);
  run<4, 8, 1, false>("blockIdx items 4  win 256", src, dst, n, st, ticket, ref);
  run<8, 1, 1, true>("blockIdx items 8  no look-back", src, dst, n, st, ticket, ref);
  run<16, 1, 1, true>("blockIdx items 16 no look-back", src, dst, n, st, ticket, ref);
  run<4, 1, 1, true>("blockIdx items 4  no look-back", src, dst, n, st, ticket, ref);
  return 0;
}
