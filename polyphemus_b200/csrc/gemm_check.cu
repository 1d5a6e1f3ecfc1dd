// Plain fp32 CUDA-core contraction used ONLY by the tests as an on-device cross-check of the tcgen05
// kernels (pb_gemm_f32_check) and the weight-operand preparation (pb_weight_prep).
#include "common.cuh"

namespace pb {

// D[m,n] = sum_k opA(m,k) * opB(k,n) (+ bias[n]);  opA = A[m*lda+k] or A[k*lda+m]; opB = B[k*ldb+n] or B[n*ldb+k]
template <int TILE, int KT>
__global__ void __launch_bounds__(256) gemm_f32_check_kernel(const float* __restrict__ a, int64_t lda,
                                                            const float* __restrict__ b, int64_t ldb,
                                                            const float* __restrict__ bias, float* __restrict__ dst,
                                                            int64_t ldd, int64_t m, int n, int k, int trans_a,
                                                            int trans_b) {
  __shared__ float sa[KT][TILE + 1];
  __shared__ float sb[KT][TILE + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * TILE;
  const int n0 = blockIdx.x * TILE;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < k; k0 += KT) {
    for (int i = threadIdx.x; i < TILE * KT; i += 256) {
      int kk, mm;
      if (trans_a) { mm = i % TILE; kk = i / TILE; } else { kk = i % KT; mm = i / KT; }
      const int64_t gm = m0 + mm;
      const int gk = k0 + kk;
      float v = 0.f;
      if (gm < m && gk < k) v = trans_a ? a[(size_t)gk * lda + gm] : a[(size_t)gm * lda + gk];
      sa[kk][mm] = v;
      int nn;
      if (trans_b) { kk = i % KT; nn = i / KT; } else { nn = i % TILE; kk = i / TILE; }
      const int gn = n0 + nn;
      const int gk2 = k0 + kk;
      v = 0.f;
      if (gn < n && gk2 < k) v = trans_b ? b[(size_t)gn * ldb + gk2] : b[(size_t)gk2 * ldb + gn];
      sb[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = sa[kk][ty * 4 + i]; bv[i] = sb[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gm = m0 + ty * 4 + i;
    if (gm >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < n) dst[(size_t)gm * ldd + gn] = acc[i][j] + (bias ? bias[gn] : 0.f);
    }
  }
}

// Wcat = [weight[0]; ...; weight[R-1]; root]  ((R+1)d x d, row-major)  and its transpose [d x (R+1)d].
template <bool BF16>
__global__ void weight_prep_kernel(const float* __restrict__ weight, const float* __restrict__ root, int n_rel, int d,
                                   void* __restrict__ w_hi, void* __restrict__ w_lo, void* __restrict__ wt_hi,
                                   void* __restrict__ wt_lo) {
  __shared__ float tile[32][33];
  const int kdim = (n_rel + 1) * d;
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  {
    const int kk = k0 + threadIdx.y, nn = n0 + threadIdx.x;
    const float v = kk < n_rel * d ? weight[(size_t)kk * d + nn] : root[(size_t)(kk - n_rel * d) * d + nn];
    tile[threadIdx.y][threadIdx.x] = v;
    const size_t o = (size_t)kk * d + nn;
    if constexpr (BF16) {
      reinterpret_cast<__nv_bfloat16*>(w_hi)[o] = __float2bfloat16_rn(v);
    } else {
      const float h = tf32_hi(v);
      reinterpret_cast<float*>(w_hi)[o] = h;
      reinterpret_cast<float*>(w_lo)[o] = v - h;
    }
  }
  __syncthreads();
  {
    const int nn = n0 + threadIdx.y, kk = k0 + threadIdx.x;
    const float v = tile[threadIdx.x][threadIdx.y];
    const size_t o = (size_t)nn * kdim + kk;
    if constexpr (BF16) {
      reinterpret_cast<__nv_bfloat16*>(wt_hi)[o] = __float2bfloat16_rn(v);
    } else {
      const float h = tf32_hi(v);
      reinterpret_cast<float*>(wt_hi)[o] = h;
      reinterpret_cast<float*>(wt_lo)[o] = v - h;
    }
  }
}

}  // namespace pb

using namespace pb;

extern "C" int pb_gemm_f32_check(const float* a, int64_t lda, const float* b, int64_t ldb, const float* bias,
                                 float* d_out, int64_t ldd, int64_t m, int32_t n, int32_t k, int32_t trans_a,
                                 int32_t trans_b, pb_stream_t stream) {
  PB_REQUIRE(a && b && d_out && m > 0 && n > 0 && k > 0, "pb_gemm_f32_check: bad arguments");
  dim3 grid((n + 63) / 64, (unsigned)((m + 63) / 64));
  gemm_f32_check_kernel<64, 16><<<grid, 256, 0, as_stream(stream)>>>(a, lda, b, ldb, bias, d_out, ldd, m, n, k, trans_a,
                                                                     trans_b);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_weight_prep(const float* weight, const float* root, int32_t n_relations, int32_t d, int32_t dtype,
                              void* wcat_hi, void* wcat_lo, void* wcat_t_hi, void* wcat_t_lo, pb_stream_t stream) {
  PB_REQUIRE(weight && root && wcat_hi && wcat_t_hi, "pb_weight_prep: null pointer");
  PB_REQUIRE(d > 0 && d % 32 == 0 && n_relations > 0, "pb_weight_prep: d must be a multiple of 32");
  PB_REQUIRE(dtype == PB_BF16 || (dtype == PB_F32 && wcat_lo && wcat_t_lo), "pb_weight_prep: PB_F32 needs lo buffers");
  dim3 grid(d / 32, (n_relations + 1) * d / 32), block(32, 32);
  if (dtype == PB_BF16)
    weight_prep_kernel<true><<<grid, block, 0, as_stream(stream)>>>(weight, root, n_relations, d, wcat_hi, wcat_lo,
                                                                    wcat_t_hi, wcat_t_lo);
  else
    weight_prep_kernel<false><<<grid, block, 0, as_stream(stream)>>>(weight, root, n_relations, d, wcat_hi, wcat_lo,
                                                                     wcat_t_hi, wcat_t_lo);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
