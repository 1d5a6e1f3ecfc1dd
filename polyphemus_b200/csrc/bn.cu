// BatchNorm (batch statistics over all node rows) + ReLU + residual, forward and backward.
//
// Replaces GCN.forward's per-layer tail (reference model.py:198-206): PyG BatchNorm == nn.BatchNorm1d(d,
// eps=1e-5, momentum=0.1) over the N node rows, F.relu, `residual + x`; and what autograd derives for them.
// All kernels are HBM-bound streaming passes with 16-byte accesses. Column reductions are two-stage with a
// fixed CTA->row-range partition and a fixed combination order (double precision in the tiny second stage),
// so results are bit-reproducible; the variance uses sums shifted by the first row to avoid cancellation.
//
// bn_coef layout (f32 [3,d]): row 0 = mean, row 1 = scale (= gamma * rstd), row 2 = beta
//   y = x_res + relu((out - mean) * scale + beta)
#include <string.h>

#include "common.cuh"

namespace pb {

// Row groups of the structured (track-relation-sorted, 128-row padded) node layout: `count[g]` valid rows starting
// at padded row `start[g]`; the rows in between are zero padding that takes no part in the statistics. A plain
// [m, d] matrix is one group {0, m}.
struct RowMap {
  int n;
  long long total;     // valid rows (= cum[n]); separate field so that kernels never index the arrays dynamically
  long long cum[5];    // cum[g] = valid rows before group g
  long long start[4];
  __device__ __forceinline__ long long padded(long long r) const {   // compact valid index -> padded row
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < n && r < cum[g + 1]) return start[g] + (r - cum[g]);
    return r;
  }
  __device__ __forceinline__ bool valid(long long p) const {         // is padded row p a real node?
    bool ok = false;
#pragma unroll
    for (int g = 0; g < 4; ++g) ok = ok || (g < n && p >= start[g] && p < start[g] + (cum[g + 1] - cum[g]));
    return ok;
  }
};

static RowMap make_rowmap(const pb_groups_t* groups, int64_t m) {
  RowMap rm;
  memset(&rm, 0, sizeof(rm));
  if (!groups || groups->n_groups <= 0) {
    rm.n = 1;
    rm.cum[1] = rm.cum[2] = rm.cum[3] = rm.cum[4] = m;
    rm.total = m;
    return rm;
  }
  rm.n = groups->n_groups;
  for (int g = 0; g < 4; ++g) {
    rm.start[g] = g < rm.n ? groups->start[g] : 0;
    rm.cum[g + 1] = rm.cum[g] + (g < rm.n ? groups->count[g] : 0);
  }
  rm.total = rm.cum[4];
  return rm;
}
static inline int64_t valid_rows(const RowMap& rm) { return rm.total; }

// Row-loop unrolling of the column passes gives each thread several independent loads in flight, but only pays if the
// registers still allow the 4 CTAs per SM the grid is sized for (592 CTAs = one wave): without the occupancy bound the
// unrolled backward passes needed a second wave and got 30 % SLOWER (197 -> 258 us); with it 197 -> 174 us.
#ifndef PB_BN_BWD_UNROLL
#define PB_BN_BWD_UNROLL 6
#endif
#ifndef PB_BN_CTAS_PER_SM
#define PB_BN_CTAS_PER_SM 4
#endif
#ifndef PB_BN_MINB
#define PB_BN_MINB 4
#endif
constexpr int kColThreads = 256;
constexpr int kColMaxCtas = 148 * PB_BN_CTAS_PER_SM;

static inline int col_ctas(int64_t m) {
  int64_t n = (m + 127) / 128;
  return (int)std::max<int64_t>(1, std::min<int64_t>(n, kColMaxCtas));
}

// Each CTA reduces a contiguous row range for every column; thread = (row group, float4 column chunk).
// partials: [gridDim.x][NV][d]
// UNROLL rows of loads in flight per thread (the cheap statistics pass takes 4; the backward passes carry too many live
// values for that and slow down).
// REVERSE: CTA b takes the row range of index gridDim.x - 1 - b (the rows the previous kernel touched last first). Tried
// for the passes that re-read a matrix right after another kernel wrote / read it (statistics after the GEMM, second
// backward pass): no measurable L2 benefit on B200 at 134 MB per matrix (44.6 vs 44.0 us, 195 vs 190 us), left off.
template <int NV, int UNROLL, bool REVERSE = false, class Load>
__device__ __forceinline__ void column_partials(const RowMap& rm, int d, float* __restrict__ partials, Load load) {
  const int64_t m = rm.total;   // valid rows; the loader receives padded row indices
  const int range = REVERSE ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x;
  extern __shared__ float red[];  // [NV][nrg][d]
  const int nchunk = d >> 2;
  const int nrg = kColThreads / nchunk;
  const int rg = threadIdx.x / nchunk, c = threadIdx.x - rg * nchunk;
  const int64_t per = (m + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)range * per;
  const int64_t r1 = r0 + per < m ? r0 + per : m;
  float4 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rg < nrg) {
    // walk the CTA's compact row range group by group: inside a group compact and padded rows differ by a constant
#pragma unroll
    for (int g = 0; g < 4; ++g) {   // groups beyond rm.n are empty (cum[g] == cum[g+1])
      const int64_t lo = r0 > rm.cum[g] ? r0 : rm.cum[g];
      const int64_t hi = r1 < rm.cum[g + 1] ? r1 : rm.cum[g + 1];
      const int64_t shift = rm.start[g] - rm.cum[g];
      // keep the row-group phase of the plain loop: thread row-group rg owns compact rows r0 + rg + i * nrg
      int64_t r = r0 + rg;
      if (r < lo) r += (lo - r + nrg - 1) / nrg * nrg;
#pragma unroll UNROLL
      for (; r < hi; r += nrg) {   // the additions stay in row order
        float4 v[NV];
        load(r + shift, c, v);
#pragma unroll
        for (int i = 0; i < NV; ++i) { acc[i].x += v[i].x; acc[i].y += v[i].y; acc[i].z += v[i].z; acc[i].w += v[i].w; }
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) reinterpret_cast<float4*>(red + ((size_t)i * nrg + rg) * d)[c] = acc[i];
  }
  __syncthreads();
  if (rg == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int g = 0; g < nrg; ++g) {
        const float4 v = reinterpret_cast<const float4*>(red + ((size_t)i * nrg + g) * d)[c];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      reinterpret_cast<float4*>(partials + ((size_t)range * NV + i) * d)[c] = s;
    }
  }
}

static inline size_t col_smem(int nv, int d) { return (size_t)nv * (kColThreads / (d / 4)) * d * sizeof(float); }

// Second stage of every column reduction: block = 32 columns x kFinLanes partial-lanes. Lane py sums partials
// py, py+kFinLanes, ... in double; the lanes are then combined in fixed order. Result in s[] of the py == 0 threads.
constexpr int kFinLanes = 32;
template <int NV>
__device__ __forceinline__ bool finalize_sums(const float* __restrict__ partials, int n_part, int d, double s[NV]) {
  __shared__ double red[NV][kFinLanes][33];
  const int ci = threadIdx.x, py = threadIdx.y;
  const int c = blockIdx.x * 32 + ci;
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  if (c < d)
    for (int p = py; p < n_part; p += kFinLanes)
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] += (double)partials[((size_t)p * NV + i) * d + c];
#pragma unroll
  for (int i = 0; i < NV; ++i) red[i][py][ci] = acc[i];
  __syncthreads();
  if (py != 0 || c >= d) return false;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double t = 0.0;
    for (int l = 0; l < kFinLanes; ++l) t += red[i][l][ci];
    s[i] = t;
  }
  return true;
}

// ------------------------------------------------------------------------------------------------ forward stats
#ifndef PB_BN_APPLY_UNROLL
#define PB_BN_APPLY_UNROLL 2
#endif
#ifndef PB_BN_APPLY_CTAS
#define PB_BN_APPLY_CTAS 32
#endif
constexpr int kApplyUnroll = PB_BN_APPLY_UNROLL;
#ifndef PB_BN_STATS_UNROLL
#define PB_BN_STATS_UNROLL 8
#endif
template <bool ABF>
__global__ void __launch_bounds__(kColThreads, PB_BN_MINB) bn_stats_partial_kernel(const void* __restrict__ out, int64_t ldo,
                                                                      const RowMap rm, int d,
                                                                      float* __restrict__ partials) {
  const size_t shift_off = (size_t)rm.padded(0) * ldo;
  column_partials<2, PB_BN_STATS_UNROLL>(rm, d, partials, [&](int64_t r, int c, float4* v) {
    const float4 x = act_ld4_stream<ABF>(out, (size_t)r * ldo + 4 * c);
    const float4 k = act_ld4<ABF>(out, shift_off + 4 * c);  // shift = first valid row (exact, cancels in the variance)
    const float4 dlt = make_float4(x.x - k.x, x.y - k.y, x.z - k.z, x.w - k.w);
    v[0] = dlt;
    v[1] = make_float4(dlt.x * dlt.x, dlt.y * dlt.y, dlt.z * dlt.z, dlt.w * dlt.w);
  });
}

__global__ void bn_stats_finalize_kernel(const float* __restrict__ partials, int n_part,
                                         const void* __restrict__ shift_row, bool abf, int64_t m, int d,
                                         const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float eps, float momentum,
                                         float* __restrict__ running_mean, float* __restrict__ running_var,
                                         float* __restrict__ save_mean_rstd, float* __restrict__ coef) {
  double s[2];
  if (!finalize_sums<2>(partials, n_part, d, s)) return;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const double s1 = s[0], s2 = s[1];
  const double n = (double)m;
  const double mean = (shift_row ? (double)act_ld1(shift_row, c, abf) : 0.0) + s1 / n;   // shift_row NULL: unshifted sums
  double var = (s2 - s1 * s1 / n) / n;
  if (var < 0.0) var = 0.0;
  const float meanf = (float)mean, varf = (float)var;
  const float rstd = (float)(1.0 / sqrt((double)varf + (double)eps));
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * meanf;
  if (running_var) {
    const float unbiased = m > 1 ? (float)(var * n / (n - 1.0)) : varf;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
  save_mean_rstd[c] = meanf;
  save_mean_rstd[d + c] = rstd;
  coef[c] = meanf;
  coef[d + c] = gamma[c] * rstd;
  coef[2 * d + c] = beta[c];
}

__global__ void bn_prepare_eval_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                       const float* __restrict__ rm, const float* __restrict__ rv, float eps, int d,
                                       float* __restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  coef[c] = rm[c];
  coef[d + c] = gamma[c] * (float)(1.0 / sqrt((double)rv[c] + (double)eps));
  coef[2 * d + c] = beta[c];
}

// ------------------------------------------------------------------------------------------------ forward apply
template <bool RELU, bool RES, bool ABF>
__global__ void __launch_bounds__(256) bn_apply_kernel(const void* __restrict__ out, int64_t ldo,
                                                      const void* __restrict__ x_res, const float* __restrict__ coef,
                                                      void* __restrict__ y, int64_t m, int d, const RowMap rm) {
  const int nchunk = d >> 2;
  const int64_t total = m * nchunk;
#pragma unroll kApplyUnroll
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / nchunk;
    const int c = (int)(i - r * nchunk);
    if (rm.n > 1 && !rm.valid(r)) {   // padding row of the structured layout stays zero
      act_st4_stream<ABF>(y, (size_t)r * d + 4 * c, make_float4(0.f, 0.f, 0.f, 0.f));
      continue;
    }
    const float4 o = act_ld4_stream<ABF>(out, (size_t)r * ldo + 4 * c);
    const float4 mu = ldg4(coef + 4 * c), sc = ldg4(coef + d + 4 * c), be = ldg4(coef + 2 * d + 4 * c);
    float4 z = make_float4((o.x - mu.x) * sc.x + be.x, (o.y - mu.y) * sc.y + be.y, (o.z - mu.z) * sc.z + be.z,
                           (o.w - mu.w) * sc.w + be.w);
    if (RELU) { z.x = fmaxf(z.x, 0.f); z.y = fmaxf(z.y, 0.f); z.z = fmaxf(z.z, 0.f); z.w = fmaxf(z.w, 0.f); }
    if (RES) {
      const float4 xr = act_ld4_stream<ABF>(x_res, (size_t)r * d + 4 * c);
      z.x += xr.x; z.y += xr.y; z.z += xr.z; z.w += xr.w;
    }
    act_st4_stream<ABF>(y, (size_t)r * d + 4 * c, z);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// pass 1: g_beta = sum gz, g_gamma = sum gz * xhat   with gz = gy * 1[(out-mean)*scale+beta > 0]
template <bool ABF>
__global__ void __launch_bounds__(kColThreads, PB_BN_MINB) bn_bwd_partial_kernel(const void* __restrict__ gy,
                                                                    const void* __restrict__ out, int64_t ldo,
                                                                    const float* __restrict__ coef,
                                                                    const float* __restrict__ mean_rstd,
                                                                    const RowMap rm, int d,
                                                                    float* __restrict__ partials) {
  column_partials<2, PB_BN_BWD_UNROLL>(rm, d, partials, [&](int64_t r, int c, float4* v) {
    const float4 g = act_ld4_stream<ABF>(gy, (size_t)r * d + 4 * c);
    const float4 o = act_ld4_stream<ABF>(out, (size_t)r * ldo + 4 * c);
    const float4 mu = ldg4(coef + 4 * c), sc = ldg4(coef + d + 4 * c), be = ldg4(coef + 2 * d + 4 * c);
    const float4 rs = ldg4(mean_rstd + d + 4 * c);
    const float4 ctr = make_float4(o.x - mu.x, o.y - mu.y, o.z - mu.z, o.w - mu.w);
    float4 gz;
    gz.x = ctr.x * sc.x + be.x > 0.f ? g.x : 0.f;
    gz.y = ctr.y * sc.y + be.y > 0.f ? g.y : 0.f;
    gz.z = ctr.z * sc.z + be.z > 0.f ? g.z : 0.f;
    gz.w = ctr.w * sc.w + be.w > 0.f ? g.w : 0.f;
    v[0] = gz;
    v[1] = make_float4(gz.x * ctr.x * rs.x, gz.y * ctr.y * rs.y, gz.z * ctr.z * rs.z, gz.w * ctr.w * rs.w);
  });
}

// sums [n_part][NV][d] -> out_v[NV][d] in double, fixed order
template <int NV>
__global__ void col_finalize_kernel(const float* __restrict__ partials, int n_part, int d, float* __restrict__ o0,
                                    float* __restrict__ o1) {
  double s[NV];
  if (!finalize_sums<NV>(partials, n_part, d, s)) return;
  const int c = blockIdx.x * 32 + threadIdx.x;
  if (o0) o0[c] = (float)s[0];
  if (NV > 1 && o1) o1[c] = (float)s[NV > 1 ? 1 : 0];
}

template <bool BF16>
__device__ __forceinline__ void store_g(void* g_hi, void* g_lo, size_t off, float4 v) {
  if constexpr (BF16) {
    st_stream2(reinterpret_cast<__nv_bfloat16*>(g_hi) + off, make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w)));
  } else {
    const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    st_stream4(reinterpret_cast<float*>(g_hi) + off, h);
    st_stream4(reinterpret_cast<float*>(g_lo) + off, make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
  }
}

// pass 2: g_out = scale * (gz - g_beta/m - xhat * g_gamma/m), written as the GEMM operand; column sums of
// g_out (the bias gradient of the layer feeding this BatchNorm) leave as partials.
template <bool BF16, bool ABF>
__global__ void __launch_bounds__(kColThreads, PB_BN_MINB) bn_bwd_apply_kernel(
    const void* __restrict__ gy, const void* __restrict__ out, int64_t ldo, const float* __restrict__ coef,
    const float* __restrict__ mean_rstd, const float* __restrict__ g_gamma, const float* __restrict__ g_beta,
    const RowMap rm, int d, void* __restrict__ g_hi, void* __restrict__ g_lo, int64_t ldg,
    float* __restrict__ partials) {
  const float inv_m = 1.f / (float)rm.total;
  column_partials<1, PB_BN_BWD_UNROLL>(rm, d, partials, [&](int64_t r, int c, float4* v) {
    const float4 g = act_ld4_stream<ABF>(gy, (size_t)r * d + 4 * c);
    const float4 o = act_ld4_stream<ABF>(out, (size_t)r * ldo + 4 * c);
    const float4 mu = ldg4(coef + 4 * c), sc = ldg4(coef + d + 4 * c), be = ldg4(coef + 2 * d + 4 * c);
    const float4 rs = ldg4(mean_rstd + d + 4 * c);
    const float4 gg = ldg4(g_gamma + 4 * c), gb = ldg4(g_beta + 4 * c);
    float4 res;
#define PB_BN_BWD(f)                                                   \
  {                                                                    \
    const float ctr = o.f - mu.f;                                      \
    const float gz = ctr * sc.f + be.f > 0.f ? g.f : 0.f;              \
    const float xhat = ctr * rs.f;                                     \
    res.f = sc.f * (gz - gb.f * inv_m - xhat * (gg.f * inv_m));        \
  }
    PB_BN_BWD(x) PB_BN_BWD(y) PB_BN_BWD(z) PB_BN_BWD(w)
#undef PB_BN_BWD
    store_g<BF16>(g_hi, g_lo, (size_t)r * ldg + 4 * c, res);
    v[0] = res;
  });
}

// zero the padding rows of the structured layout in the GEMM-operand copy of the gradient (the weight-gradient GEMM
// contracts over ALL rows): rows [start[g] + count[g], start[g+1]) of every group, at most 127 each
struct PadRows { int n; long long begin[4], end[4]; };
template <bool BF16>
__global__ void zero_pad_rows_kernel(void* __restrict__ g_hi, void* __restrict__ g_lo, int64_t ldg, int d, const PadRows pr) {
  const int nchunk = d >> 2;
  for (int g = 0; g < pr.n; ++g) {
    const long long rows = pr.end[g] - pr.begin[g];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * nchunk; i += (long long)gridDim.x * blockDim.x) {
      const long long r = pr.begin[g] + i / nchunk;
      const int c = (int)(i % nchunk);
      store_g<BF16>(g_hi, g_lo, (size_t)r * ldg + 4 * c, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
}

template <bool BF16>
__global__ void __launch_bounds__(kColThreads) grad_prep_kernel(const float* __restrict__ g, int64_t ldg_in, int64_t m,
                                                               int d, void* __restrict__ g_hi, void* __restrict__ g_lo,
                                                               int64_t ldg, float* __restrict__ partials) {
  RowMap rm;
  rm.n = 1; rm.total = m; rm.cum[0] = 0; rm.cum[1] = rm.cum[2] = rm.cum[3] = rm.cum[4] = m;
  rm.start[0] = rm.start[1] = rm.start[2] = rm.start[3] = 0;
  column_partials<1, 1>(rm, d, partials, [&](int64_t r, int c, float4* v) {
    const float4 x = ld_stream4(g + (size_t)r * ldg_in + 4 * c);
    store_g<BF16>(g_hi, g_lo, (size_t)r * ldg + 4 * c, x);
    v[0] = x;
  });
}

static int check_md(int64_t m, int d, const char* who) {
  PB_REQUIRE(m > 0, "%s: m must be positive", who);
  PB_REQUIRE(d >= 64 && d % 64 == 0 && d <= 1024, "%s: d=%d must be a multiple of 64 in [64, 1024]", who, d);
  return PB_OK;
}

}  // namespace pb

using namespace pb;

extern "C" size_t pb_bn_workspace_bytes(int64_t m, int32_t d) {
  if (m <= 0 || d <= 0) return 0;
  return align_up((size_t)col_ctas(m) * 2 * d * sizeof(float), 256);
}

static int check_act(int32_t act_dtype, int64_t ld, const char* who) {
  PB_REQUIRE(act_dtype == PB_F32 || act_dtype == PB_BF16, "%s: act_dtype %d", who, act_dtype);
  PB_REQUIRE(ld % 4 == 0, "%s: row stride must be a multiple of 4 elements", who);
  return PB_OK;
}

extern "C" int pb_bn_stats(const void* out, int64_t ldo, int64_t m, int32_t d, const pb_groups_t* groups,
                           const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                           float* running_var, float* save_mean_rstd, float* bn_coef, void* workspace,
                           size_t workspace_bytes, int32_t act_dtype, pb_stream_t stream) {
  if (int rca = check_act(act_dtype, ldo, "pb_bn_stats")) return rca;
  const bool abf = act_dtype == PB_BF16;
  const size_t esz = abf ? 2 : 4;
  int rc = check_md(m, d, "pb_bn_stats");
  if (rc) return rc;
  PB_REQUIRE(out && gamma && beta && save_mean_rstd && bn_coef && workspace, "pb_bn_stats: null pointer");
  PB_REQUIRE(ldo >= d && ldo % 4 == 0, "pb_bn_stats: bad ldo");
  PB_REQUIRE(workspace_bytes >= pb_bn_workspace_bytes(m, d), "pb_bn_stats: workspace too small");
  cudaStream_t st = as_stream(stream);
  const RowMap rm = make_rowmap(groups, m);
  const int64_t mv = valid_rows(rm);
  PB_REQUIRE(mv > 0 && mv <= m, "pb_bn_stats: bad row groups");
  const int ctas = col_ctas(mv);
  float* partials = reinterpret_cast<float*>(workspace);
  if (abf) bn_stats_partial_kernel<true><<<ctas, kColThreads, col_smem(2, d), st>>>(out, ldo, rm, d, partials);
  else bn_stats_partial_kernel<false><<<ctas, kColThreads, col_smem(2, d), st>>>(out, ldo, rm, d, partials);
  PB_LAUNCH_CHECK();
  const char* base = static_cast<const char*>(out);
  const void* shift_row = base + (size_t)(rm.n > 0 ? rm.start[0] + 0 : 0) * ldo * esz;   // == padded(0) when group 0 is non-empty
  for (int g = 0; g < rm.n; ++g)
    if (rm.cum[g + 1] > rm.cum[g]) { shift_row = base + (size_t)rm.start[g] * ldo * esz; break; }
  bn_stats_finalize_kernel<<<(d + 31) / 32, dim3(32, kFinLanes), 0, st>>>(partials, ctas, shift_row, abf, mv, d, gamma, beta, eps,
                                                                           momentum, running_mean, running_var,
                                                                           save_mean_rstd, bn_coef);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

// Statistics from column partials [n_partials][2][d] (sum, sum of squares; unshifted) that another kernel left behind —
// the forward GEMM's epilogue (pb_rgcn_gemm_fwd_bn) — instead of a second pass over `out`.
extern "C" int pb_bn_finalize(const float* partials, int64_t n_partials, int64_t m_valid, int32_t d, const float* gamma,
                              const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                              float* save_mean_rstd, float* bn_coef, pb_stream_t stream) {
  PB_REQUIRE(partials && gamma && beta && save_mean_rstd && bn_coef, "pb_bn_finalize: null pointer");
  PB_REQUIRE(n_partials > 0 && n_partials < ((int64_t)1 << 31) && m_valid > 0 && d > 0, "pb_bn_finalize: bad sizes");
  bn_stats_finalize_kernel<<<(d + 31) / 32, dim3(32, kFinLanes), 0, as_stream(stream)>>>(
      partials, (int)n_partials, nullptr, false, m_valid, d, gamma, beta, eps, momentum, running_mean, running_var,
      save_mean_rstd, bn_coef);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_bn_prepare_eval(const float* gamma, const float* beta, const float* running_mean,
                                  const float* running_var, float eps, int32_t d, float* bn_coef, pb_stream_t stream) {
  PB_REQUIRE(gamma && beta && running_mean && running_var && bn_coef && d > 0, "pb_bn_prepare_eval: bad arguments");
  bn_prepare_eval_kernel<<<(d + 127) / 128, 128, 0, as_stream(stream)>>>(gamma, beta, running_mean, running_var, eps, d,
                                                                        bn_coef);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_bn_relu_res_fwd(const void* out, int64_t ldo, const void* x_res, const float* bn_coef, void* y,
                                  int64_t m, int32_t d, const pb_groups_t* groups, int32_t apply_relu,
                                  int32_t act_dtype, pb_stream_t stream) {
  if (int rca = check_act(act_dtype, ldo, "pb_bn_relu_res_fwd")) return rca;
  int rc = check_md(m, d, "pb_bn_relu_res_fwd");
  if (rc) return rc;
  PB_REQUIRE(out && bn_coef && y, "pb_bn_relu_res_fwd: null pointer");
  PB_REQUIRE(ldo >= d && ldo % 4 == 0, "pb_bn_relu_res_fwd: bad ldo");
  const int64_t total = m * (d / 4);
  const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * PB_BN_APPLY_CTAS);
  cudaStream_t st = as_stream(stream);
  const RowMap rm = make_rowmap(groups, m);
#define PB_BN_APPLY(RELU, RES, ABF) \
  bn_apply_kernel<RELU, RES, ABF><<<grid, 256, 0, st>>>(out, ldo, x_res, bn_coef, y, m, d, rm)
#define PB_BN_APPLY_ACT(RELU, RES) \
  do { if (act_dtype == PB_BF16) PB_BN_APPLY(RELU, RES, true); else PB_BN_APPLY(RELU, RES, false); } while (0)
  if (apply_relu) {
    if (x_res) PB_BN_APPLY_ACT(true, true); else PB_BN_APPLY_ACT(true, false);
  } else {
    if (x_res) PB_BN_APPLY_ACT(false, true); else PB_BN_APPLY_ACT(false, false);
  }
#undef PB_BN_APPLY_ACT
#undef PB_BN_APPLY
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_bn_relu_res_bwd(const void* gy, const void* out, int64_t ldo, const float* gamma,
                                  const float* save_mean_rstd, const float* bn_coef, int64_t m, int32_t d,
                                  const pb_groups_t* groups, int32_t dtype, void* g_hi, void* g_lo, int64_t ldg,
                                  float* g_gamma, float* g_beta, float* g_bias, void* workspace,
                                  size_t workspace_bytes, int32_t act_dtype, pb_stream_t stream) {
  if (int rca = check_act(act_dtype, ldo, "pb_bn_relu_res_bwd")) return rca;
  const bool abf = act_dtype == PB_BF16;
  int rc = check_md(m, d, "pb_bn_relu_res_bwd");
  if (rc) return rc;
  (void)gamma;
  PB_REQUIRE(gy && out && save_mean_rstd && bn_coef && g_hi && g_gamma && g_beta && workspace,
             "pb_bn_relu_res_bwd: null pointer");
  PB_REQUIRE(dtype == PB_BF16 || (dtype == PB_F32 && g_lo), "pb_bn_relu_res_bwd: PB_F32 needs g_lo");
  PB_REQUIRE(ldo >= d && ldo % 4 == 0 && ldg >= d && ldg % 8 == 0, "pb_bn_relu_res_bwd: bad leading dimension");
  PB_REQUIRE(workspace_bytes >= pb_bn_workspace_bytes(m, d), "pb_bn_relu_res_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const RowMap rm = make_rowmap(groups, m);
  const int64_t mv = valid_rows(rm);
  PB_REQUIRE(mv > 0 && mv <= m, "pb_bn_relu_res_bwd: bad row groups");
  const int ctas = col_ctas(mv);
  float* partials = reinterpret_cast<float*>(workspace);
  if (abf) bn_bwd_partial_kernel<true><<<ctas, kColThreads, col_smem(2, d), st>>>(gy, out, ldo, bn_coef, save_mean_rstd, rm, d, partials);
  else bn_bwd_partial_kernel<false><<<ctas, kColThreads, col_smem(2, d), st>>>(gy, out, ldo, bn_coef, save_mean_rstd, rm, d, partials);
  PB_LAUNCH_CHECK();
  col_finalize_kernel<2><<<(d + 31) / 32, dim3(32, kFinLanes), 0, st>>>(partials, ctas, d, g_beta, g_gamma);
  PB_LAUNCH_CHECK();
#define PB_BN_BWD_APPLY(BF, ABF)                                                                                     \
  bn_bwd_apply_kernel<BF, ABF><<<ctas, kColThreads, col_smem(1, d), st>>>(gy, out, ldo, bn_coef, save_mean_rstd, g_gamma, \
                                                                          g_beta, rm, d, g_hi, g_lo, ldg, partials)
  if (dtype == PB_BF16) { if (abf) PB_BN_BWD_APPLY(true, true); else PB_BN_BWD_APPLY(true, false); }
  else { if (abf) PB_BN_BWD_APPLY(false, true); else PB_BN_BWD_APPLY(false, false); }
#undef PB_BN_BWD_APPLY
  PB_LAUNCH_CHECK();
  if (g_bias) {
    col_finalize_kernel<1><<<(d + 31) / 32, dim3(32, kFinLanes), 0, st>>>(partials, ctas, d, g_bias, nullptr);
    PB_LAUNCH_CHECK();
  }
  if (groups && groups->n_groups > 0) {     // padding rows of the operand copy: written here, not pre-zeroed by the caller
    PadRows pr;
    memset(&pr, 0, sizeof(pr));
    long long n_pad = 0;
    for (int g = 0; g < rm.n; ++g) {
      const long long b = rm.start[g] + (rm.cum[g + 1] - rm.cum[g]);
      const long long e = g + 1 < rm.n ? rm.start[g + 1] : (long long)m;
      if (e > b) { pr.begin[pr.n] = b; pr.end[pr.n] = e; ++pr.n; n_pad += e - b; }
    }
    if (n_pad > 0) {
      const unsigned grid = (unsigned)std::min<long long>((n_pad * (d / 4) + 255) / 256, 64);
      if (dtype == PB_BF16) zero_pad_rows_kernel<true><<<grid, 256, 0, st>>>(g_hi, g_lo, ldg, d, pr);
      else zero_pad_rows_kernel<false><<<grid, 256, 0, st>>>(g_hi, g_lo, ldg, d, pr);
      PB_LAUNCH_CHECK();
    }
  }
  return PB_OK;
}

extern "C" int pb_grad_prep(const float* g, int64_t ldg_in, int64_t m, int32_t d, int32_t dtype, void* g_hi, void* g_lo,
                            int64_t ldg, float* g_bias, void* workspace, size_t workspace_bytes, pb_stream_t stream) {
  int rc = check_md(m, d, "pb_grad_prep");
  if (rc) return rc;
  PB_REQUIRE(g && g_hi && workspace, "pb_grad_prep: null pointer");
  PB_REQUIRE(dtype == PB_BF16 || (dtype == PB_F32 && g_lo), "pb_grad_prep: PB_F32 needs g_lo");
  PB_REQUIRE(ldg_in >= d && ldg_in % 4 == 0 && ldg >= d && ldg % 8 == 0, "pb_grad_prep: bad leading dimension");
  PB_REQUIRE(workspace_bytes >= pb_bn_workspace_bytes(m, d), "pb_grad_prep: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int ctas = col_ctas(m);
  float* partials = reinterpret_cast<float*>(workspace);
  if (dtype == PB_BF16)
    grad_prep_kernel<true><<<ctas, kColThreads, col_smem(1, d), st>>>(g, ldg_in, m, d, g_hi, g_lo, ldg, partials);
  else
    grad_prep_kernel<false><<<ctas, kColThreads, col_smem(1, d), st>>>(g, ldg_in, m, d, g_hi, g_lo, ldg, partials);
  PB_LAUNCH_CHECK();
  if (g_bias) {
    col_finalize_kernel<1><<<(d + 31) / 32, dim3(32, kFinLanes), 0, st>>>(partials, ctas, d, g_bias, nullptr);
    PB_LAUNCH_CHECK();
  }
  return PB_OK;
}
