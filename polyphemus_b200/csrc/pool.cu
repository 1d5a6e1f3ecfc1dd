// Per-bar segment operators around the message-passing stacks. The nodes of one bar are contiguous in the batched
// graph (bar_ptr[b] .. bar_ptr[b+1], from pb_graph_count), so both directions are plain segmented loops — one warp per
// bar, fixed summation order, no atomics (the index_add_ formulation is neither deterministic nor cheap).
//
//   pool   : PyG GlobalAttention (model.py:335-340, 409):  alpha = softmax_bar(gate),  out[b] = sum_v alpha_v h[v]
//   expand : x[v] = z[bar(v)]  (model.py:542-546: every node starts from its bar's code) and its segment-sum gradient
#include "common.cuh"

namespace pb {
namespace {

constexpr int kThreads = 256;
constexpr uint32_t kFull = 0xffffffffu;
constexpr int kMaxSeg = 128;   // 4 tracks x 32 timesteps

__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

template <int CPL>
__global__ void __launch_bounds__(kThreads)
bar_pool_fwd_kernel(const float* __restrict__ h, long ldh, const float* __restrict__ gate,
                    const int* __restrict__ bar_ptr, int n_bars, int d, float* __restrict__ alpha,
                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  if (b >= n_bars) return;
  const int beg = __ldg(bar_ptr + b), end = __ldg(bar_ptr + b + 1);
  const int nchunk = d >> 2;
  float gv[kMaxSeg / 32];
  float m = -INFINITY;
#pragma unroll
  for (int q = 0; q < kMaxSeg / 32; ++q) {
    const int v = beg + lane + 32 * q;
    gv[q] = v < end ? __ldg(gate + v) : -INFINITY;
    m = fmaxf(m, gv[q]);
  }
  m = wmax(m);
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < kMaxSeg / 32; ++q) {
    gv[q] = beg + lane + 32 * q < end ? expf(gv[q] - m) : 0.f;
    s += gv[q];
  }
  s = wsum(s);
  const float inv = 1.0f / (s + 1e-16f);
#pragma unroll
  for (int q = 0; q < kMaxSeg / 32; ++q) {
    gv[q] *= inv;
    const int v = beg + lane + 32 * q;
    if (v < end) alpha[v] = gv[q];
  }
  float4 acc[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int q = 0; q < kMaxSeg / 32; ++q) {
    const int base = beg + 32 * q;
    if (base >= end) break;
    const int cnt = min(32, end - base);
#pragma unroll 4
    for (int i = 0; i < cnt; ++i) {
      const float a = __shfl_sync(kFull, gv[q], i);
      const float* row = h + static_cast<long>(base + i) * ldh;
#pragma unroll
      for (int j = 0; j < CPL; ++j)
        if (lane + 32 * j < nchunk) {
          const float4 x = ld_stream4(row + 4 * (lane + 32 * j));
          acc[j].x += a * x.x; acc[j].y += a * x.y; acc[j].z += a * x.z; acc[j].w += a * x.w;
        }
    }
  }
#pragma unroll
  for (int j = 0; j < CPL; ++j)
    if (lane + 32 * j < nchunk) *reinterpret_cast<float4*>(out + static_cast<long>(b) * d + 4 * (lane + 32 * j)) = acc[j];
}

// g_h[v] = alpha_v g_out[b];  g_gate[v] = alpha_v (<g_out[b], h[v]> - sum_u alpha_u <g_out[b], h[u]>)
template <int CPL>
__global__ void __launch_bounds__(kThreads)
bar_pool_bwd_kernel(const float* __restrict__ h, long ldh, const float* __restrict__ alpha,
                    const int* __restrict__ bar_ptr, int n_bars, int d, const float* __restrict__ g_out,
                    float* __restrict__ g_h, long ldgh, float* __restrict__ g_gate) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  if (b >= n_bars) return;
  const int beg = __ldg(bar_ptr + b), end = __ldg(bar_ptr + b + 1);
  const int nchunk = d >> 2;
  float4 go[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j)
    go[j] = lane + 32 * j < nchunk ? ldg4(g_out + static_cast<long>(b) * d + 4 * (lane + 32 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
  float av[kMaxSeg / 32], dot[kMaxSeg / 32];
#pragma unroll
  for (int q = 0; q < kMaxSeg / 32; ++q) {
    const int v = beg + lane + 32 * q;
    av[q] = v < end ? __ldg(alpha + v) : 0.f;
    dot[q] = 0.f;
  }
  float mean = 0.f;
#pragma unroll
  for (int q = 0; q < kMaxSeg / 32; ++q) {
    const int base = beg + 32 * q;
    if (base >= end) break;
    const int cnt = min(32, end - base);
#pragma unroll 2
    for (int i = 0; i < cnt; ++i) {
      const float a = __shfl_sync(kFull, av[q], i);
      const float* row = h + static_cast<long>(base + i) * ldh;
      float* grow = g_h + static_cast<long>(base + i) * ldgh;
      float p = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j)
        if (lane + 32 * j < nchunk) {
          const float4 x = ld_stream4(row + 4 * (lane + 32 * j));
          p += go[j].x * x.x + go[j].y * x.y + go[j].z * x.z + go[j].w * x.w;
          st_stream4(grow + 4 * (lane + 32 * j), make_float4(a * go[j].x, a * go[j].y, a * go[j].z, a * go[j].w));
        }
      p = wsum(p);
      if (lane == i) dot[q] = p;
      mean += a * p;
    }
  }
#pragma unroll
  for (int q = 0; q < kMaxSeg / 32; ++q) {
    const int v = beg + lane + 32 * q;
    if (v < end) g_gate[v] = av[q] * (dot[q] - mean);
  }
}

template <int CPL>
__global__ void __launch_bounds__(kThreads)
bar_expand_fwd_kernel(const float* __restrict__ z, const int* __restrict__ bar_ptr, int n_bars, int d,
                      float* __restrict__ x, long ldx) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  if (b >= n_bars) return;
  const int beg = __ldg(bar_ptr + b), end = __ldg(bar_ptr + b + 1);
  const int nchunk = d >> 2;
  float4 zv[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j)
    zv[j] = lane + 32 * j < nchunk ? ldg4(z + static_cast<long>(b) * d + 4 * (lane + 32 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int v = beg; v < end; ++v)
#pragma unroll
    for (int j = 0; j < CPL; ++j)
      if (lane + 32 * j < nchunk) *reinterpret_cast<float4*>(x + static_cast<long>(v) * ldx + 4 * (lane + 32 * j)) = zv[j];
}

template <int CPL>
__global__ void __launch_bounds__(kThreads)
bar_expand_bwd_kernel(const float* __restrict__ g_x, long ldg, const int* __restrict__ bar_ptr, int n_bars, int d,
                      float* __restrict__ g_z) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  if (b >= n_bars) return;
  const int beg = __ldg(bar_ptr + b), end = __ldg(bar_ptr + b + 1);
  const int nchunk = d >> 2;
  float4 acc[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int v = beg; v < end; ++v)
#pragma unroll
    for (int j = 0; j < CPL; ++j)
      if (lane + 32 * j < nchunk) {
        const float4 g = ld_stream4(g_x + static_cast<long>(v) * ldg + 4 * (lane + 32 * j));
        acc[j].x += g.x; acc[j].y += g.y; acc[j].z += g.z; acc[j].w += g.w;
      }
#pragma unroll
  for (int j = 0; j < CPL; ++j)
    if (lane + 32 * j < nchunk) *reinterpret_cast<float4*>(g_z + static_cast<long>(b) * d + 4 * (lane + 32 * j)) = acc[j];
}

int bar_check(const char* who, const int32_t* bar_ptr, int64_t n_bars, int32_t d) {
  PB_REQUIRE(bar_ptr && n_bars > 0 && n_bars < (1ll << 26), "%s: bar segments", who);
  PB_REQUIRE(d >= 4 && d % 4 == 0 && d <= 1024, "%s: d=%d must be a multiple of 4, at most 1024", who, d);
  return PB_OK;
}

}  // namespace
}  // namespace pb

using namespace pb;

#define PB_BAR_DISPATCH(KERNEL, ...)                                                        \
  do {                                                                                      \
    const unsigned grid = static_cast<unsigned>((n_bars + 7) / 8);                          \
    const int cpl = (d / 4 + 31) / 32;                                                      \
    if (cpl <= 1) KERNEL<1><<<grid, kThreads, 0, as_stream(stream)>>>(__VA_ARGS__);         \
    else if (cpl <= 2) KERNEL<2><<<grid, kThreads, 0, as_stream(stream)>>>(__VA_ARGS__);    \
    else if (cpl <= 4) KERNEL<4><<<grid, kThreads, 0, as_stream(stream)>>>(__VA_ARGS__);    \
    else KERNEL<8><<<grid, kThreads, 0, as_stream(stream)>>>(__VA_ARGS__);                  \
    PB_LAUNCH_CHECK();                                                                      \
  } while (0)

extern "C" int pb_bar_pool_fwd(const float* h, int64_t ldh, const float* gate, const int32_t* bar_ptr, int64_t n_bars,
                               int32_t d, float* alpha, float* out, pb_stream_t stream) {
  if (int rc = bar_check("pb_bar_pool_fwd", bar_ptr, n_bars, d)) return rc;
  PB_REQUIRE(h && gate && alpha && out && ldh >= d && ldh % 4 == 0, "pb_bar_pool_fwd: bad arguments");
  PB_BAR_DISPATCH(bar_pool_fwd_kernel, h, ldh, gate, bar_ptr, static_cast<int>(n_bars), d, alpha, out);
  return PB_OK;
}

extern "C" int pb_bar_pool_bwd(const float* h, int64_t ldh, const float* alpha, const int32_t* bar_ptr, int64_t n_bars,
                               int32_t d, const float* g_out, float* g_h, int64_t ldgh, float* g_gate,
                               pb_stream_t stream) {
  if (int rc = bar_check("pb_bar_pool_bwd", bar_ptr, n_bars, d)) return rc;
  PB_REQUIRE(h && alpha && g_out && g_h && g_gate && ldh >= d && ldh % 4 == 0 && ldgh >= d && ldgh % 4 == 0,
             "pb_bar_pool_bwd: bad arguments");
  PB_BAR_DISPATCH(bar_pool_bwd_kernel, h, ldh, alpha, bar_ptr, static_cast<int>(n_bars), d, g_out, g_h, ldgh, g_gate);
  return PB_OK;
}

extern "C" int pb_bar_expand_fwd(const float* z, const int32_t* bar_ptr, int64_t n_bars, int32_t d, float* x,
                                 int64_t ldx, pb_stream_t stream) {
  if (int rc = bar_check("pb_bar_expand_fwd", bar_ptr, n_bars, d)) return rc;
  PB_REQUIRE(z && x && ldx >= d && ldx % 4 == 0, "pb_bar_expand_fwd: bad arguments");
  PB_BAR_DISPATCH(bar_expand_fwd_kernel, z, bar_ptr, static_cast<int>(n_bars), d, x, ldx);
  return PB_OK;
}

extern "C" int pb_bar_expand_bwd(const float* g_x, int64_t ldg, const int32_t* bar_ptr, int64_t n_bars, int32_t d,
                                 float* g_z, pb_stream_t stream) {
  if (int rc = bar_check("pb_bar_expand_bwd", bar_ptr, n_bars, d)) return rc;
  PB_REQUIRE(g_x && g_z && ldg >= d && ldg % 4 == 0, "pb_bar_expand_bwd: bad arguments");
  PB_BAR_DISPATCH(bar_expand_bwd_kernel, g_x, ldg, bar_ptr, static_cast<int>(n_bars), d, g_z);
  return PB_OK;
}
