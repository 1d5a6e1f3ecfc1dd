// Message + mean aggregation of the relational graph convolution, forward and backward.
//
// Replaces, per layer and per relation, the reference's sequence (SURVEY.md §2.3 steps d-j):
//   index_select(x, src)  ->  nn.Linear(32,d)(one-hot dist)  ->  mul  ->  relu  ->  dropout(0.1)
//   ->  scatter_add by dst + count + divide            (GCL.message model.py:123-135, propagate model.py:110)
// with ONE pass over a destination-sorted CSR: each warp owns a destination node, lanes own channels
// (16-byte vector loads), every (dst, relation) segment is summed in edge order in registers and divided
// by its length — no atomics, no E x d intermediates, bit-reproducible. The Linear on a one-hot input is a
// 32-row table lookup T[dist] (pb_edge_table_fwd).
//
// The kernel writes the tensor-core operand A = [H_0 | ... | H_{R-1} | x] directly in the GEMM's dtype.
#include "common.cuh"

namespace pb {

// ---------------------------------------------------------------------------------------------- edge table
__global__ void edge_table_fwd_kernel(const float* __restrict__ w, const float* __restrict__ b, int d,
                                      float* __restrict__ table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // i = k*d + c
  if (i >= PB_N_DISTS * d) return;
  const int k = i / d, c = i - k * d;
  table[i] = w[c * PB_N_DISTS + k] + b[c];
}

// partials [P][32][d] -> g_w [d][32], g_b [d]; fixed summation order (p ascending, then k ascending)
__global__ void __launch_bounds__(1024) edge_table_bwd_kernel(const float* __restrict__ partials, int n_partials,
                                                              int d, float* __restrict__ g_w,
                                                              float* __restrict__ g_b) {
  __shared__ float tile[32][33];
  const int ci = threadIdx.x, k = threadIdx.y;
  const int c = blockIdx.x * 32 + ci;
  float s = 0.f;
  if (c < d) {
    const float* p = partials + (size_t)k * d + c;
    const size_t stride = (size_t)PB_N_DISTS * d;
    for (int i = 0; i < n_partials; ++i) s += p[i * stride];
    g_w[(size_t)c * PB_N_DISTS + k] = s;
  }
  tile[k][ci] = s;
  __syncthreads();
  if (k == 0 && c < d) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) t += tile[j][ci];
    g_b[c] = t;
  }
}

// ---------------------------------------------------------------------------------------------- forward
template <bool BF16>
__device__ __forceinline__ void store_operand(void* a_hi, void* a_lo, size_t elem_off, float4 v) {
  if constexpr (BF16) {
    uint2 p = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    st_stream2(reinterpret_cast<__nv_bfloat16*>(a_hi) + elem_off, p);
  } else {
    float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
    st_stream4(reinterpret_cast<float*>(a_hi) + elem_off, h);
    st_stream4(reinterpret_cast<float*>(a_lo) + elem_off, l);
  }
}

// CPL = float4 chunks per lane (d <= 128*CPL). One warp per destination node.
template <bool BF16, bool DROPOUT, int CPL>
__global__ void __launch_bounds__(256) agg_fwd_kernel(const int* __restrict__ in_ptr, const int* __restrict__ in_edge,
                                                      const int* __restrict__ in_eid, const float* __restrict__ x,
                                                      const float* __restrict__ table, void* __restrict__ a_hi,
                                                      void* __restrict__ a_lo, int64_t lda, int64_t n_nodes, int d,
                                                      int n_rel, uint32_t thresh, float keep_scale, uint64_t seed) {
  const int lane = threadIdx.x & 31;
  const int nchunk = d >> 2;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < n_nodes; v += n_warps) {
    const size_t row = (size_t)v * lda;
    // root block: the node's own features, converted to the operand dtype
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int c = lane + 32 * j;
      if (c < nchunk) store_operand<BF16>(a_hi, a_lo, row + (size_t)n_rel * d + 4 * c, ldg4(x + (size_t)v * d + 4 * c));
    }
    const int* seg = in_ptr + v * n_rel;
    int beg = __ldg(seg);
    for (int r = 0; r < n_rel; ++r) {
      const int end = __ldg(seg + r + 1);
      float4 acc[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int e = beg; e < end; ++e) {
        const uint32_t pk = (uint32_t)__ldg(in_edge + e);
        const size_t src = pk & 0x03FFFFFFu;
        const int dist = pk >> 26;
        uint32_t eid = 0;
        if constexpr (DROPOUT) eid = (uint32_t)__ldg(in_eid + e);
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int c = lane + 32 * j;
          if (c < nchunk) {
            const float4 xs = ldg4(x + src * d + 4 * c);
            const float4 t = ldg4(table + (size_t)dist * d + 4 * c);
            float4 m = make_float4(fmaxf(xs.x * t.x, 0.f), fmaxf(xs.y * t.y, 0.f), fmaxf(xs.z * t.z, 0.f),
                                   fmaxf(xs.w * t.w, 0.f));
            if constexpr (DROPOUT) {
              bool keep[4];
              dropout_keep4(seed, eid, (uint32_t)c, thresh, keep);
              m.x = keep[0] ? m.x * keep_scale : 0.f;
              m.y = keep[1] ? m.y * keep_scale : 0.f;
              m.z = keep[2] ? m.z * keep_scale : 0.f;
              m.w = keep[3] ? m.w * keep_scale : 0.f;
            }
            acc[j].x += m.x; acc[j].y += m.y; acc[j].z += m.z; acc[j].w += m.w;
          }
        }
      }
      const int cnt = end - beg;
      const float fc = (float)(cnt > 1 ? cnt : 1);
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int c = lane + 32 * j;
        if (c < nchunk) {
          float4 h = acc[j];
          if (cnt > 1) { h.x = h.x / fc; h.y = h.y / fc; h.z = h.z / fc; h.w = h.w / fc; }  // scatter-mean
          store_operand<BF16>(a_hi, a_lo, row + (size_t)r * d + 4 * c, h);
        }
      }
      beg = end;
    }
  }
}

// ---------------------------------------------------------------------------------------------- backward
constexpr int kBwdThreads = 128;
constexpr int kBwdPartials = 444;  // 3 CTAs on each of 148 SMs; fixed so the reduction order never changes

template <bool BF16>
__device__ __forceinline__ float4 load_grad4(const void* d_a, size_t elem_off) {
  if constexpr (BF16) {
    const uint2 p = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(d_a) + elem_off));
    const float2 a = unpack_bf16x2(p.x), b = unpack_bf16x2(p.y);
    return make_float4(a.x, a.y, b.x, b.y);
  } else {
    return ldg4(reinterpret_cast<const float*>(d_a) + elem_off);
  }
}

// Scatter-by-source without atomics: each node u gathers the gradients of the segments its out-edges feed.
// A CTA owns a contiguous node range; thread t owns channel chunks (fixed), so the per-distance table
// gradient accumulates in shared memory with a fixed order per address -> deterministic.
//   TPN  threads per node (d/4 capped at 128), G = 128/TPN nodes processed concurrently per CTA
//   CPT  float4 chunks per thread (d/512 when d > 512)
template <bool BF16, bool DROPOUT, int CPT>
__global__ void __launch_bounds__(kBwdThreads) agg_bwd_kernel(
    const int* __restrict__ out_ptr, const int4* __restrict__ out_rec, const float* __restrict__ x,
    const float* __restrict__ table, const void* __restrict__ d_a, int64_t ldda, const float* __restrict__ gy_res,
    float* __restrict__ gx, float* __restrict__ dt_partials, int64_t n_nodes, int d, int n_rel, int tpn,
    uint32_t thresh, float keep_scale, uint64_t seed) {
  extern __shared__ float dts[];  // [G][32][d]
  const int groups = kBwdThreads / tpn;
  const int g = threadIdx.x / tpn, tc = threadIdx.x % tpn;
  const int nchunk = d >> 2;
  for (int i = threadIdx.x; i < groups * PB_N_DISTS * d; i += kBwdThreads) dts[i] = 0.f;
  __syncthreads();
  float* my_dt = dts + (size_t)g * PB_N_DISTS * d;

  const int64_t per_cta = (n_nodes + gridDim.x - 1) / gridDim.x;
  const int64_t first = (int64_t)blockIdx.x * per_cta;
  const int64_t last = first + per_cta < n_nodes ? first + per_cta : n_nodes;
  for (int64_t u = first + g; u < last; u += groups) {
    float4 xu[CPT], acc[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = tc + tpn * j;
      if (c < nchunk) {
        xu[j] = ldg4(x + (size_t)u * d + 4 * c);
        acc[j] = load_grad4<BF16>(d_a, (size_t)u * ldda + (size_t)n_rel * d + 4 * c);  // root branch
        if (gy_res) {
          const float4 r = ldg4(gy_res + (size_t)u * d + 4 * c);  // residual branch
          acc[j].x += r.x; acc[j].y += r.y; acc[j].z += r.z; acc[j].w += r.w;
        }
      }
    }
    const int beg = __ldg(out_ptr + u), end = __ldg(out_ptr + u + 1);
    for (int i = beg; i < end; ++i) {
      const int4 rec = __ldg(out_rec + i);
      const int rel = rec.y & 0xff, dist = rec.y >> 8;
      const float inv_scale = keep_scale;  // 1/(1-p) when dropout is on, else 1
      const float fc = (float)rec.w;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = tc + tpn * j;
        if (c < nchunk) {
          const float4 dh = load_grad4<BF16>(d_a, (size_t)rec.x * ldda + (size_t)rel * d + 4 * c);
          const float4 t = ldg4(table + (size_t)dist * d + 4 * c);
          bool keep[4] = {true, true, true, true};
          if constexpr (DROPOUT) dropout_keep4(seed, (uint32_t)rec.z, (uint32_t)c, thresh, keep);
          float4 ds;
          ds.x = (keep[0] && xu[j].x * t.x > 0.f) ? (rec.w > 1 ? dh.x / fc : dh.x) : 0.f;
          ds.y = (keep[1] && xu[j].y * t.y > 0.f) ? (rec.w > 1 ? dh.y / fc : dh.y) : 0.f;
          ds.z = (keep[2] && xu[j].z * t.z > 0.f) ? (rec.w > 1 ? dh.z / fc : dh.z) : 0.f;
          ds.w = (keep[3] && xu[j].w * t.w > 0.f) ? (rec.w > 1 ? dh.w / fc : dh.w) : 0.f;
          if constexpr (DROPOUT) { ds.x *= inv_scale; ds.y *= inv_scale; ds.z *= inv_scale; ds.w *= inv_scale; }
          acc[j].x += ds.x * t.x; acc[j].y += ds.y * t.y; acc[j].z += ds.z * t.z; acc[j].w += ds.w * t.w;
          float4* slot = reinterpret_cast<float4*>(my_dt + (size_t)dist * d + 4 * c);
          float4 cur = *slot;
          cur.x += ds.x * xu[j].x; cur.y += ds.y * xu[j].y; cur.z += ds.z * xu[j].z; cur.w += ds.w * xu[j].w;
          *slot = cur;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = tc + tpn * j;
      if (c < nchunk) st_stream4(gx + (size_t)u * d + 4 * c, acc[j]);
    }
  }
  __syncthreads();
  float* out = dt_partials + (size_t)blockIdx.x * PB_N_DISTS * d;
  for (int i = threadIdx.x; i < PB_N_DISTS * d; i += kBwdThreads) {
    float s = 0.f;
    for (int gg = 0; gg < groups; ++gg) s += dts[(size_t)gg * PB_N_DISTS * d + i];
    out[i] = s;
  }
}

__global__ void dropout_mask_kernel(int64_t n_edges, int d, uint32_t thresh, uint64_t seed, uint8_t* __restrict__ keep) {
  const int nchunk = d >> 2;
  const int64_t total = n_edges * nchunk;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / nchunk;
    const int c = (int)(i - e * nchunk);
    bool k[4];
    dropout_keep4(seed, (uint32_t)e, (uint32_t)c, thresh, k);
    uchar4 o = make_uchar4(k[0], k[1], k[2], k[3]);
    reinterpret_cast<uchar4*>(keep)[i] = o;
  }
}

static int check_csr(const pb_csr_t* g, int d, const char* who) {
  PB_REQUIRE(g && g->in_ptr && g->in_edge && g->out_ptr && g->out_rec, "%s: incomplete CSR plan", who);
  PB_REQUIRE(g->n_nodes > 0 && g->n_relations > 0, "%s: empty graph", who);
  PB_REQUIRE(d >= 64 && d % 64 == 0 && d <= 1024, "%s: d=%d must be a multiple of 64 in [64, 1024]", who, d);
  return PB_OK;
}

}  // namespace pb

using namespace pb;

extern "C" int pb_edge_table_fwd(const float* nn_weight, const float* nn_bias, int32_t d, float* table,
                                 pb_stream_t stream) {
  PB_REQUIRE(nn_weight && nn_bias && table && d > 0, "pb_edge_table_fwd: bad arguments");
  const int total = PB_N_DISTS * d;
  edge_table_fwd_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(nn_weight, nn_bias, d, table);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_edge_table_bwd(const float* dtable_partials, int32_t n_partials, int32_t d, float* g_nn_weight,
                                 float* g_nn_bias, pb_stream_t stream) {
  PB_REQUIRE(dtable_partials && g_nn_weight && g_nn_bias && n_partials > 0 && d > 0, "pb_edge_table_bwd: bad arguments");
  edge_table_bwd_kernel<<<(d + 31) / 32, dim3(32, 32), 0, as_stream(stream)>>>(dtable_partials, n_partials, d,
                                                                               g_nn_weight, g_nn_bias);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

template <bool BF16, bool DROP>
static int launch_agg_fwd(const pb_csr_t* g, const float* x, int d, const float* table, void* a_hi, void* a_lo,
                          int64_t lda, uint32_t thresh, float scale, uint64_t seed, cudaStream_t st) {
  const int cpl = (d + 127) / 128;
  const int threads = 256;
  const int64_t want = (g->n_nodes * 32 + threads - 1) / threads;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count() * 8 * 4));
#define PB_AGG_FWD(CPL)                                                                                         \
  agg_fwd_kernel<BF16, DROP, CPL><<<grid, threads, 0, st>>>(g->in_ptr, g->in_edge, g->in_eid, x, table, a_hi,    \
                                                            a_lo, lda, g->n_nodes, d, g->n_relations, thresh,    \
                                                            scale, seed)
  if (cpl <= 1) PB_AGG_FWD(1);
  else if (cpl <= 2) PB_AGG_FWD(2);
  else if (cpl <= 4) PB_AGG_FWD(4);
  else PB_AGG_FWD(8);
#undef PB_AGG_FWD
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_agg_fwd(const pb_csr_t* csr, const float* x, int32_t d, const float* table, void* a_hi, void* a_lo,
                          int64_t lda, int32_t dtype, float p_drop, uint64_t seed, pb_stream_t stream) {
  int rc = check_csr(csr, d, "pb_agg_fwd");
  if (rc) return rc;
  PB_REQUIRE(x && table && a_hi, "pb_agg_fwd: null pointer");
  PB_REQUIRE(dtype == PB_BF16 || (dtype == PB_F32 && a_lo), "pb_agg_fwd: PB_F32 needs a_lo");
  PB_REQUIRE(lda >= (int64_t)(csr->n_relations + 1) * d && lda % 8 == 0, "pb_agg_fwd: bad lda");
  PB_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "pb_agg_fwd: p_drop out of range");
  PB_REQUIRE(p_drop == 0.f || csr->in_eid, "pb_agg_fwd: dropout needs in_eid");
  cudaStream_t st = as_stream(stream);
  const uint32_t thresh = dropout_thresh(p_drop);
  const float scale = 1.f / (1.f - p_drop);
  if (dtype == PB_BF16)
    return p_drop > 0.f ? launch_agg_fwd<true, true>(csr, x, d, table, a_hi, a_lo, lda, thresh, scale, seed, st)
                        : launch_agg_fwd<true, false>(csr, x, d, table, a_hi, a_lo, lda, thresh, scale, seed, st);
  return p_drop > 0.f ? launch_agg_fwd<false, true>(csr, x, d, table, a_hi, a_lo, lda, thresh, scale, seed, st)
                      : launch_agg_fwd<false, false>(csr, x, d, table, a_hi, a_lo, lda, thresh, scale, seed, st);
}

extern "C" int32_t pb_agg_bwd_num_partials(void) { return kBwdPartials; }

template <bool BF16, bool DROP>
static int launch_agg_bwd(const pb_csr_t* g, const float* x, int d, const float* table, const void* d_a, int64_t ldda,
                          const float* gy_res, float* gx, float* dtp, uint32_t thresh, float scale, uint64_t seed,
                          cudaStream_t st) {
  const int tpn = std::min(kBwdThreads, d / 4);
  const int groups = kBwdThreads / tpn;
  const int cpt = (d / 4 + tpn - 1) / tpn;
  const size_t smem = (size_t)groups * PB_N_DISTS * d * sizeof(float);
#define PB_AGG_BWD(CPT)                                                                                          \
  do {                                                                                                           \
    PB_CUDA(cudaFuncSetAttribute(agg_bwd_kernel<BF16, DROP, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                 (int)smem));                                                                    \
    agg_bwd_kernel<BF16, DROP, CPT><<<kBwdPartials, kBwdThreads, smem, st>>>(                                    \
        g->out_ptr, reinterpret_cast<const int4*>(g->out_rec), x, table, d_a, ldda, gy_res, gx, dtp, g->n_nodes, \
        d, g->n_relations, tpn, thresh, scale, seed);                                                            \
  } while (0)
  if (cpt <= 1) PB_AGG_BWD(1);
  else PB_AGG_BWD(2);
#undef PB_AGG_BWD
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_agg_bwd(const pb_csr_t* csr, const float* x, int32_t d, const float* table, const void* d_a,
                          int64_t ldda, int32_t dtype, const float* gy_res, float* gx, float* dtable_partials,
                          float p_drop, uint64_t seed, pb_stream_t stream) {
  int rc = check_csr(csr, d, "pb_agg_bwd");
  if (rc) return rc;
  PB_REQUIRE(x && table && d_a && gx && dtable_partials, "pb_agg_bwd: null pointer");
  PB_REQUIRE(dtype == PB_BF16 || dtype == PB_F32, "pb_agg_bwd: bad dtype");
  PB_REQUIRE(ldda >= (int64_t)(csr->n_relations + 1) * d && ldda % 8 == 0, "pb_agg_bwd: bad ldda");
  PB_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "pb_agg_bwd: p_drop out of range");
  cudaStream_t st = as_stream(stream);
  const uint32_t thresh = dropout_thresh(p_drop);
  const float scale = 1.f / (1.f - p_drop);
  if (dtype == PB_BF16)
    return p_drop > 0.f ? launch_agg_bwd<true, true>(csr, x, d, table, d_a, ldda, gy_res, gx, dtable_partials, thresh, scale, seed, st)
                        : launch_agg_bwd<true, false>(csr, x, d, table, d_a, ldda, gy_res, gx, dtable_partials, thresh, scale, seed, st);
  return p_drop > 0.f ? launch_agg_bwd<false, true>(csr, x, d, table, d_a, ldda, gy_res, gx, dtable_partials, thresh, scale, seed, st)
                      : launch_agg_bwd<false, false>(csr, x, d, table, d_a, ldda, gy_res, gx, dtable_partials, thresh, scale, seed, st);
}

extern "C" int pb_dropout_mask(int64_t n_edges, int32_t d, float p_drop, uint64_t seed, uint8_t* keep,
                               pb_stream_t stream) {
  PB_REQUIRE(keep && n_edges >= 0 && d > 0 && d % 4 == 0, "pb_dropout_mask: bad arguments");
  if (n_edges == 0) return PB_OK;
  const int64_t total = n_edges * (d / 4);
  const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
  dropout_mask_kernel<<<grid, 256, 0, as_stream(stream)>>>(n_edges, d, dropout_thresh(p_drop), seed, keep);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
