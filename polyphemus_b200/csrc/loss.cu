// Row-wise masked cross entropy over the padded un-embedding heads (training.py:100-101, 320-330:
// nn.CrossEntropyLoss(ignore_index=PAD) on the pitch / duration logits of every token slot).
//
// The logits of one training step are [N*15, 192|128] — gigabytes — and the library formulation touches them six
// times (fp32 copy, log_softmax, gather, and the three backward passes with a zero-filled scatter). Here each is read
// once per direction: forward keeps only nll[r] and the row's log-sum-exp, backward rebuilds softmax from the stored
// log-sum-exp and writes the logit gradient in the logits' own type. HBM-bound, one warp per RPW rows, all loads of
// a warp issued before the first use.
#include "common.cuh"

namespace pb {
namespace {

constexpr int kCeThreads = 256;

template <bool BF16, int CHUNKS>
__device__ __forceinline__ void ce_load_row(const void* logits, long ld, long r, int classes, int lane, bool live,
                                            float (&v)[CHUNKS][BF16 ? 8 : 4]) {
  constexpr int V = BF16 ? 8 : 4;
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const int col = (c * 32 + lane) * V;
    if (live && col < classes) {
      if constexpr (BF16) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(logits) + r * ld + col));
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[c][2 * j] = __uint_as_float(w[j] << 16);
          v[c][2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
        }
      } else {
        const float4 raw = ldg4(static_cast<const float*>(logits) + r * ld + col);
        v[c][0] = raw.x, v[c][1] = raw.y, v[c][2] = raw.z, v[c][3] = raw.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) v[c][j] = -INFINITY;
    }
  }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <bool BF16, int CHUNKS, int RPW>
__global__ void __launch_bounds__(kCeThreads)
ce_fwd_kernel(const void* __restrict__ logits, long ld, long rows, int classes, const int* __restrict__ target,
              int ignore, float* __restrict__ nll, float* __restrict__ lse) {
  constexpr int V = BF16 ? 8 : 4;
  const int lane = threadIdx.x & 31;
  const long r0 = ((static_cast<long>(blockIdx.x) * kCeThreads + threadIdx.x) >> 5) * RPW;
  float v[RPW][CHUNKS][V];
  int tgt[RPW];
#pragma unroll
  for (int i = 0; i < RPW; ++i) tgt[i] = (r0 + i < rows) ? __ldg(target + r0 + i) : ignore;
#pragma unroll
  for (int i = 0; i < RPW; ++i) ce_load_row<BF16, CHUNKS>(logits, ld, r0 + i, classes, lane, tgt[i] != ignore, v[i]);
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const long r = r0 + i;
    if (tgt[i] == ignore) {                       // warp-uniform
      if (lane == 0 && r < rows) nll[r] = 0.f, lse[r] = 0.f;
      continue;
    }
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
      for (int j = 0; j < V; ++j) m = fmaxf(m, v[i][c][j]);
    m = warp_max(m);
    float s = 0.f, xt = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
      for (int j = 0; j < V; ++j) {
        s += expf(v[i][c][j] - m);
        if ((c * 32 + lane) * V + j == tgt[i]) xt = v[i][c][j];
      }
    s = warp_sum(s);
    xt = warp_sum(xt);
    if (lane == 0) {
      const float l = m + logf(s);
      lse[r] = l;
      nll[r] = l - xt;
    }
  }
}

template <bool BF16, int CHUNKS, int RPW>
__global__ void __launch_bounds__(kCeThreads)
ce_bwd_kernel(const void* __restrict__ logits, long ld, long rows, int classes, const int* __restrict__ target,
              int ignore, const float* __restrict__ lse, const float* __restrict__ row_grad, void* __restrict__ grad,
              long ldg) {
  constexpr int V = BF16 ? 8 : 4;
  const int lane = threadIdx.x & 31;
  const long r0 = ((static_cast<long>(blockIdx.x) * kCeThreads + threadIdx.x) >> 5) * RPW;
  float v[RPW][CHUNKS][V];
  int tgt[RPW];
  float l[RPW], rg[RPW];
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const bool in = r0 + i < rows;
    tgt[i] = in ? __ldg(target + r0 + i) : ignore;
    l[i] = in ? __ldg(lse + r0 + i) : 0.f;
    rg[i] = in ? __ldg(row_grad + r0 + i) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < RPW; ++i) ce_load_row<BF16, CHUNKS>(logits, ld, r0 + i, classes, lane, tgt[i] != ignore, v[i]);
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const long r = r0 + i;
    if (r >= rows) break;
    const bool live = tgt[i] != ignore;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const int col = (c * 32 + lane) * V;
      if (col >= classes) continue;
      float o[V];
#pragma unroll
      for (int j = 0; j < V; ++j)
        o[j] = live ? (expf(v[i][c][j] - l[i]) - (col + j == tgt[i] ? 1.f : 0.f)) * rg[i] : 0.f;
      if constexpr (BF16) {
        uint4 pk;
        uint32_t* w = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __nv_bfloat162 b = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
          w[j] = *reinterpret_cast<const uint32_t*>(&b);
        }
        *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(grad) + r * ldg + col) = pk;
      } else {
        *reinterpret_cast<float4*>(static_cast<float*>(grad) + r * ldg + col) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

int ce_check(const void* logits, int64_t ld, int32_t dtype, int64_t rows, int32_t classes, int* chunks) {
  PB_REQUIRE(dtype == PB_BF16 || dtype == PB_F32, "pb_ce: dtype %d", dtype);
  const int v = dtype == PB_BF16 ? 8 : 4;
  PB_REQUIRE(rows >= 0 && classes > 0 && classes % v == 0 && ld % v == 0 && ld >= classes,
             "pb_ce: classes %d / ld %lld must be multiples of %d", classes, static_cast<long long>(ld), v);
  PB_REQUIRE(reinterpret_cast<uintptr_t>(logits) % 16 == 0, "pb_ce: logits must be 16-byte aligned");
  const int need = (classes + 32 * v - 1) / (32 * v);
  PB_REQUIRE(need <= 4, "pb_ce: at most %d classes", 4 * 32 * v);
  *chunks = need <= 1 ? 1 : need <= 2 ? 2 : 4;
  return PB_OK;
}

}  // namespace
}  // namespace pb

using namespace pb;

#define PB_CE_DISPATCH(KERNEL, ...)                                                                  \
  do {                                                                                               \
    if (dtype == PB_BF16) {                                                                          \
      const unsigned grid = static_cast<unsigned>(((rows + 3) / 4 + 7) / 8);                         \
      if (chunks == 1) KERNEL<true, 1, 4><<<grid, kCeThreads, 0, as_stream(stream)>>>(__VA_ARGS__);  \
      else if (chunks == 2) KERNEL<true, 2, 4><<<grid, kCeThreads, 0, as_stream(stream)>>>(__VA_ARGS__); \
      else KERNEL<true, 4, 4><<<grid, kCeThreads, 0, as_stream(stream)>>>(__VA_ARGS__);              \
    } else {                                                                                         \
      const unsigned grid = static_cast<unsigned>(((rows + 1) / 2 + 7) / 8);                         \
      if (chunks == 1) KERNEL<false, 1, 2><<<grid, kCeThreads, 0, as_stream(stream)>>>(__VA_ARGS__); \
      else if (chunks == 2) KERNEL<false, 2, 2><<<grid, kCeThreads, 0, as_stream(stream)>>>(__VA_ARGS__); \
      else KERNEL<false, 4, 2><<<grid, kCeThreads, 0, as_stream(stream)>>>(__VA_ARGS__);             \
    }                                                                                                \
  } while (0)

extern "C" int pb_ce_fwd(const void* logits, int64_t ld, int32_t dtype, int64_t rows, int32_t classes,
                         const int32_t* target, int32_t ignore_index, float* nll, float* lse, pb_stream_t stream) {
  int chunks = 0;
  if (int rc = ce_check(logits, ld, dtype, rows, classes, &chunks)) return rc;
  if (rows == 0) return PB_OK;
  PB_CE_DISPATCH(ce_fwd_kernel, logits, ld, rows, classes, target, ignore_index, nll, lse);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_ce_bwd(const void* logits, int64_t ld, int32_t dtype, int64_t rows, int32_t classes,
                         const int32_t* target, int32_t ignore_index, const float* lse, const float* row_grad,
                         void* grad, int64_t ldg, pb_stream_t stream) {
  int chunks = 0;
  if (int rc = ce_check(logits, ld, dtype, rows, classes, &chunks)) return rc;
  PB_REQUIRE(ldg % (dtype == PB_BF16 ? 8 : 4) == 0 && ldg >= classes && reinterpret_cast<uintptr_t>(grad) % 16 == 0,
             "pb_ce_bwd: grad stride %lld / alignment", static_cast<long long>(ldg));
  if (rows == 0) return PB_OK;
  PB_CE_DISPATCH(ce_bwd_kernel, logits, ld, rows, classes, target, ignore_index, lse, row_grad, grad, ldg);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
