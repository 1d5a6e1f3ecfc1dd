// Deterministic two-level exclusive prefix sum over int32 (integer adds: order independent).
#pragma once
#include "common.cuh"

namespace pb {

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanChunk = kScanThreads * kScanItems;

__device__ __forceinline__ int warp_incl_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// exclusive scan across a 1024-thread block; returns exclusive prefix of `v`, block total in `total`
__device__ __forceinline__ int block_excl_scan(int v, int& total) {
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = warp_incl_scan(v);
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = warp_sums[lane];
    int wi = warp_incl_scan(w);
    warp_sums[lane] = wi - w;   // exclusive warp offsets
  }
  __syncthreads();
  int excl = incl - v + warp_sums[wid];
  // block total = exclusive offset of last warp + its inclusive sum
  __shared__ int s_total;
  if (threadIdx.x == kScanThreads - 1) s_total = excl + v;
  __syncthreads();
  total = s_total;
  __syncthreads();
  return excl;
}

static __global__ void __launch_bounds__(kScanThreads) scan_block_sums_kernel(const int* __restrict__ in, int64_t n,
                                                                      int* __restrict__ block_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) s += in[base + i];
  int total;
  block_excl_scan(s, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

static __global__ void __launch_bounds__(kScanThreads) scan_mid_kernel(int* __restrict__ block_sums, int64_t nb) {
  int carry = 0;
  for (int64_t start = 0; start < nb; start += kScanThreads) {
    int64_t i = start + threadIdx.x;
    int v = i < nb ? block_sums[i] : 0;
    int total;
    int excl = block_excl_scan(v, total);
    if (i < nb) block_sums[i] = carry + excl;
    carry += total;
  }
}

static __global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int* __restrict__ in, int64_t n,
                                                                 const int* __restrict__ block_offs,
                                                                 int* __restrict__ out) {
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = base + i < n ? in[base + i] : 0;
    s += v[i];
  }
  int total;
  int excl = block_excl_scan(s, total) + block_offs[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = excl;
    excl += v[i];
    if (base + i == n - 1) out[n] = excl;   // total in the extra slot
  }
}

static inline size_t scan_workspace_bytes(int64_t n) {
  int64_t nb = (n + kScanChunk - 1) / kScanChunk;
  return align_up((size_t)(nb > 0 ? nb : 1) * sizeof(int), 256);
}

// out has n+1 entries; `in` and `out` may NOT alias. ws: scan_workspace_bytes(n).
static inline int exclusive_scan_i32(const int* in, int* out, int64_t n, void* ws, cudaStream_t st) {
  if (n <= 0) {
    PB_CUDA(cudaMemsetAsync(out, 0, sizeof(int), st));
    return PB_OK;
  }
  int64_t nb = (n + kScanChunk - 1) / kScanChunk;
  int* block_sums = reinterpret_cast<int*>(ws);
  scan_block_sums_kernel<<<(unsigned)nb, kScanThreads, 0, st>>>(in, n, block_sums);
  PB_LAUNCH_CHECK();
  scan_mid_kernel<<<1, kScanThreads, 0, st>>>(block_sums, nb);
  PB_LAUNCH_CHECK();
  scan_apply_kernel<<<(unsigned)nb, kScanThreads, 0, st>>>(in, n, block_sums, out);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

}  // namespace pb
