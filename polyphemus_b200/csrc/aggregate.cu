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

// partials [PB_DIST_ITEMS][d] (one row per work item, items of distance k = [item_ptr[k], item_ptr[k+1])) ->
// g_w [d][32], g_b [d]. Grid (d/32, 32 distances); block = 32 channels x 8 item-lanes: lane j adds items
// i0+j, i0+j+8, ... and the 8 lanes are combined in fixed order -> deterministic. g_b is finished by a second
// tiny kernel that adds the 32 distances of g_w in order.
__global__ void __launch_bounds__(256) edge_table_bwd_kernel(const float* __restrict__ partials,
                                                             const int* __restrict__ item_ptr, int d,
                                                             float* __restrict__ g_w) {
  __shared__ float red[8][33];
  const int ci = threadIdx.x, j = threadIdx.y, k = blockIdx.y;
  const int c = blockIdx.x * 32 + ci;
  const int i0 = __ldg(item_ptr + k), i1 = __ldg(item_ptr + k + 1);
  float s = 0.f;
  if (c < d)
    for (int i = i0 + j; i < i1; i += 8) s += partials[(size_t)i * d + c];
  red[j][ci] = s;
  __syncthreads();
  if (j == 0 && c < d) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][ci];
    g_w[(size_t)c * PB_N_DISTS + k] = t;
  }
}

__global__ void edge_table_bias_kernel(const float* __restrict__ g_w, int d, float* __restrict__ g_b) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  float t = 0.f;
#pragma unroll
  for (int k = 0; k < PB_N_DISTS; ++k) t += g_w[(size_t)c * PB_N_DISTS + k];
  g_b[c] = t;
}

// ---------------------------------------------------------------------------------------------- dropout bits
// GCL.message's dropout (model.py:133). The keep decision of every (edge, channel) is drawn once per layer call
// by a counter-based hash (common.cuh) and packed 1 bit per channel; forward, operand recompute and backward all
// read the same bits, so the hot kernels carry no RNG arithmetic.
// Layout: u16 word [(e * G16 + jj) * 32 + l] holds chunks c = l + 32 * (4 jj + nib), nib = 0..3, 4 bits each
// (chunk = 4 consecutive channels) — i.e. exactly the channels lane l of a warp owns in agg_fwd.
static inline int keep_words(int d) { return ((d + 511) / 512) * 32; }   // u16 words per edge

// The four hashes of a thread share the edge and differ by 32 in the chunk index, so their SplitMix64 inputs
// seed + GOLDEN * (ctr + 1), ctr = (eid << 32 | chunk), are  base + k * (32 * GOLDEN)  with one base per thread: the
// counter multiply is done once (and only its 32-bit pieces: GOLDEN * (eid << 32) keeps just the low word of
// GOLDEN_lo * eid). Same values as common.cuh::dropout_bits — the integer pipe is what bounds this kernel.
template <int G16, bool EXACT>   // EXACT: d == 512 * G16, every chunk of the word exists
__global__ void __launch_bounds__(256) dropout_bits_kernel(int64_t n_edges, int d, uint32_t thresh16, uint64_t seed,
                                                           uint16_t* __restrict__ bits) {
  constexpr uint64_t kGolden = 0x9E3779B97F4A7C15ull;
  const int nchunk = d >> 2;
  const int64_t total = n_edges * G16 * 32;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(i & 31);
    const int64_t q = i >> 5;
    const int jj = G16 == 1 ? 0 : (int)(q & (G16 - 1));
    const uint32_t e = (uint32_t)(G16 == 1 ? q : q / G16);
    const uint32_t c0 = (uint32_t)(l + 128 * jj);                       // chunk of nibble 0; nibble k is c0 + 32 k
    // seed + GOLDEN * ((e << 32) + c0 + 1)
    uint64_t z0 = seed + ((uint64_t)((uint32_t)kGolden * e) << 32) + kGolden * (uint64_t)(c0 + 1u);
    uint32_t w = 0;
#pragma unroll
    for (int nib = 0; nib < 4; ++nib) {
      if (EXACT || (int)c0 + 32 * nib < nchunk) {
        const uint64_t r = splitmix64_mix(z0);
        const uint32_t lo = (uint32_t)r, hi = (uint32_t)(r >> 32);
        const uint32_t k4 = (uint32_t)((lo & 0xFFFFu) >= thresh16) | ((uint32_t)((lo >> 16) >= thresh16) << 1) |
                            ((uint32_t)((hi & 0xFFFFu) >= thresh16) << 2) | ((uint32_t)((hi >> 16) >= thresh16) << 3);
        w |= k4 << (4 * nib);
      }
      z0 += 32ull * kGolden;
    }
    bits[i] = (uint16_t)w;
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

// 4 keep bits of chunk c of edge eid
__device__ __forceinline__ uint32_t keep_nibble(const uint16_t* __restrict__ bits, int g16, uint32_t eid, int c) {
  const int j = c >> 5;
  const uint32_t w = __ldg(bits + ((size_t)eid * g16 + (j >> 2)) * 32 + (c & 31));
  return (w >> (4 * (j & 3))) & 0xFu;
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

// CPL = float4 chunks per lane; EXACT: d == 128 * CPL (no tail predicates). One warp per destination node. The
// node's R+1 segment offsets and its (<= 32 at a time) edge records are fetched by one coalesced load each and
// broadcast with shuffles, so the only dependent round trip before the row gathers is that single record load.
template <bool BF16, bool DROPOUT, int CPL, bool EXACT, bool ABF>
__global__ void __launch_bounds__(256) agg_fwd_kernel(const int* __restrict__ in_ptr, const int* __restrict__ in_edge,
                                                      const int* __restrict__ in_eid, const void* __restrict__ x,
                                                      const float* __restrict__ table, void* __restrict__ a_hi,
                                                      void* __restrict__ a_lo, int64_t lda, int64_t n_nodes, int d,
                                                      int n_rel, const uint16_t* __restrict__ keep_bits, float keep_scale,
                                                      const int* __restrict__ node_order) {
  constexpr uint32_t kFull = 0xffffffffu;
  constexpr int G16 = (CPL + 3) / 4;
  const int lane = threadIdx.x & 31;
  const int nchunk = d >> 2;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t vi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; vi < n_nodes; vi += n_warps) {
    const int64_t v = node_order ? (int64_t)__ldg(node_order + vi) : vi;
    const size_t row = (size_t)v * lda;
    const int my_ptr = lane <= n_rel ? __ldg(in_ptr + v * n_rel + lane) : 0;
    const int beg_all = __shfl_sync(kFull, my_ptr, 0);
    const int end_all = __shfl_sync(kFull, my_ptr, n_rel);
    int base = beg_all;
    uint32_t my_pk = base + lane < end_all ? (uint32_t)__ldg(in_edge + base + lane) : 0u;
    uint32_t my_eid = 0;
    if constexpr (DROPOUT) my_eid = base + lane < end_all ? (uint32_t)__ldg(in_eid + base + lane) : 0u;
    // root block: the node's own features, converted to the operand dtype
    const size_t xrow = (size_t)v * d + 4 * lane;
#pragma unroll
    for (int j = 0; j < CPL; ++j)
      if (EXACT || lane + 32 * j < nchunk)
        store_operand<BF16>(a_hi, a_lo, row + (size_t)n_rel * d + 4 * (lane + 32 * j), act_ld4<ABF>(x, xrow + 128 * j));
    int e = beg_all;
    for (int r = 0; r < n_rel; ++r) {
      const int seg_end = __shfl_sync(kFull, my_ptr, r + 1);
      const int cnt = seg_end - e;
      float4 acc[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (; e < seg_end; ++e) {
        if (e - base == 32) {  // warp-uniform: next batch of records (degree > 32 only)
          base = e;
          my_pk = base + lane < end_all ? (uint32_t)__ldg(in_edge + base + lane) : 0u;
          if constexpr (DROPOUT) my_eid = base + lane < end_all ? (uint32_t)__ldg(in_eid + base + lane) : 0u;
        }
        const uint32_t pk = __shfl_sync(kFull, my_pk, e - base);
        const size_t srow = (size_t)(pk & 0x03FFFFFFu) * d + 4 * lane;
        const float* trow = table + (size_t)(pk >> 26) * d + 4 * lane;
        uint32_t kw[G16];
        if constexpr (DROPOUT) {
          const uint32_t eid = __shfl_sync(kFull, my_eid, e - base);
#pragma unroll
          for (int q = 0; q < G16; ++q) kw[q] = __ldg(keep_bits + ((size_t)eid * G16 + q) * 32 + lane);
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          if (EXACT || lane + 32 * j < nchunk) {
            const float4 xs = act_ld4<ABF>(x, srow + 128 * j);
            const float4 t = ldg4(trow + 128 * j);
            float4 m = make_float4(fmaxf(xs.x * t.x, 0.f), fmaxf(xs.y * t.y, 0.f), fmaxf(xs.z * t.z, 0.f),
                                   fmaxf(xs.w * t.w, 0.f));
            if constexpr (DROPOUT) {
              const uint32_t nib = kw[j >> 2] >> (4 * (j & 3));
              m.x = (nib & 1u) ? m.x * keep_scale : 0.f;
              m.y = (nib & 2u) ? m.y * keep_scale : 0.f;
              m.z = (nib & 4u) ? m.z * keep_scale : 0.f;
              m.w = (nib & 8u) ? m.w * keep_scale : 0.f;
            }
            acc[j].x += m.x; acc[j].y += m.y; acc[j].z += m.z; acc[j].w += m.w;
          }
        }
      }
      // scatter-mean: one reciprocal per segment (<= 1 ulp from the reference's division)
      const float inv = cnt > 1 ? 1.0f / (float)cnt : 1.0f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        if (EXACT || lane + 32 * j < nchunk) {
          const float4 h = make_float4(acc[j].x * inv, acc[j].y * inv, acc[j].z * inv, acc[j].w * inv);
          store_operand<BF16>(a_hi, a_lo, row + (size_t)r * d + 4 * (lane + 32 * j), h);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- backward
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
template <bool BF16> struct GradRaw { using type = float4; };
template <> struct GradRaw<true> { using type = uint2; };
template <bool BF16>
__device__ __forceinline__ typename GradRaw<BF16>::type load_grad_raw(const void* d_a, size_t elem_off) {
  if constexpr (BF16)
    return __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(d_a) + elem_off));
  else
    return ldg4(reinterpret_cast<const float*>(d_a) + elem_off);
}
template <bool BF16>
__device__ __forceinline__ float4 unpack_grad_raw(typename GradRaw<BF16>::type p) {
  if constexpr (BF16) {
    const float2 a = unpack_bf16x2(p.x), b = unpack_bf16x2(p.y);
    return make_float4(a.x, a.y, b.x, b.y);
  } else {
    return p;
  }
}
template <bool BF16>
__device__ __forceinline__ void store_q(void* q, size_t elem_off, float4 v) {
  if constexpr (BF16)
    st_stream2(reinterpret_cast<__nv_bfloat16*>(q) + elem_off, make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w)));
  else
    st_stream4(reinterpret_cast<float*>(q) + elem_off, v);
}

// (1) Scatter-by-source without atomics: one warp per source node u gathers the gradients of the segments its
// out-edges feed (same shape as the forward: coalesced record load, shuffles, 16-byte row gathers, full occupancy)
// and emits, per out-edge position, the row q_e = ds_e * x[u] that the edge-table gradient needs.
template <bool BF16, bool DROPOUT, int CPL, bool EXACT, bool ABF>
__global__ void __launch_bounds__(256, (BF16 && CPL <= 4) ? 3 : 1) agg_bwd_dx_kernel(
    const int* __restrict__ out_ptr, const int4* __restrict__ out_rec, const void* __restrict__ x,
    const float* __restrict__ table, const void* __restrict__ d_a, int64_t ldda, const void* __restrict__ gy_res,
    void* __restrict__ gx, void* __restrict__ q_buf, int64_t n_nodes, int d, int n_rel,
    const uint16_t* __restrict__ keep_bits, float keep_scale, const int* __restrict__ node_order) {
  constexpr uint32_t kFull = 0xffffffffu;
  constexpr int G16 = (CPL + 3) / 4;
  const int lane = threadIdx.x & 31;
  const int nchunk = d >> 2;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t ui = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ui < n_nodes; ui += n_warps) {
    const int64_t u = node_order ? (int64_t)__ldg(node_order + ui) : ui;
    const int my_ptr = lane < 2 ? __ldg(out_ptr + u + lane) : 0;
    const int beg = __shfl_sync(kFull, my_ptr, 0), end = __shfl_sync(kFull, my_ptr, 1);
    int base = beg;
    int4 my_rec = base + lane < end ? __ldg(out_rec + base + lane) : make_int4(0, 0, 0, 0);
    float4 xu[CPL], acc[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      if (EXACT || lane + 32 * j < nchunk) {
        const size_t c4 = 4 * (size_t)(lane + 32 * j);
        xu[j] = act_ld4<ABF>(x, (size_t)u * d + c4);
        acc[j] = load_grad4<BF16>(d_a, (size_t)u * ldda + (size_t)n_rel * d + c4);   // root branch
        if (gy_res) {
          const float4 r = act_ld4_stream<ABF>(gy_res, (size_t)u * d + c4);           // residual branch
          acc[j].x += r.x; acc[j].y += r.y; acc[j].z += r.z; acc[j].w += r.w;
        }
      }
    }
    // software pipeline over the out-edges: the gradient row and keep-bits of edge i+1 are in flight while edge i
    // is consumed (a warp has ~3.5 out-edges; without this each one is a full DRAM round trip)
    typename GradRaw<BF16>::type nraw[CPL];
    uint32_t nkw[G16];
    int nmeta = 0, ncnt = 0;
    auto fetch = [&](int i) {
      if (i - base == 32) {  // warp-uniform (out-degree > 32 only)
        base = i;
        my_rec = base + lane < end ? __ldg(out_rec + base + lane) : make_int4(0, 0, 0, 0);
      }
      const int dst = __shfl_sync(kFull, my_rec.x, i - base);
      nmeta = __shfl_sync(kFull, my_rec.y, i - base);
      ncnt = __shfl_sync(kFull, my_rec.w, i - base);
      const size_t goff = (size_t)dst * ldda + (size_t)(nmeta & 0xff) * d + 4 * lane;
#pragma unroll
      for (int j = 0; j < CPL; ++j)
        if (EXACT || lane + 32 * j < nchunk) nraw[j] = load_grad_raw<BF16>(d_a, goff + 128 * j);
      if constexpr (DROPOUT) {
        const uint32_t eid = (uint32_t)__shfl_sync(kFull, my_rec.z, i - base);
#pragma unroll
        for (int q = 0; q < G16; ++q) nkw[q] = __ldg(keep_bits + ((size_t)eid * G16 + q) * 32 + lane);
      }
    };
    if (beg < end) fetch(beg);
    for (int i = beg; i < end; ++i) {
      typename GradRaw<BF16>::type raw[CPL];
      uint32_t kw[G16];
#pragma unroll
      for (int j = 0; j < CPL; ++j) raw[j] = nraw[j];
#pragma unroll
      for (int q = 0; q < G16; ++q) kw[q] = nkw[q];
      const int meta = nmeta, cnt = ncnt;
      if (i + 1 < end) fetch(i + 1);
      const float* trow = table + (size_t)(meta >> 8) * d + 4 * lane;
      // d(mean)/d(sum) = 1/|segment| (one reciprocal per edge), times the dropout scale
      float coef = cnt > 1 ? 1.0f / (float)cnt : 1.0f;
      if constexpr (DROPOUT) coef *= keep_scale;
      const size_t qrow = (size_t)i * d + 4 * lane;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        if (EXACT || lane + 32 * j < nchunk) {
          float4 ds = unpack_grad_raw<BF16>(raw[j]);
          const float4 t = ldg4(trow + 128 * j);
          const float4 xv = xu[j];
          uint32_t nib = 0xFu;
          if constexpr (DROPOUT) nib = kw[j >> 2] >> (4 * (j & 3));
          ds.x = ((nib & 1u) && xv.x * t.x > 0.f) ? ds.x * coef : 0.f;
          ds.y = ((nib & 2u) && xv.y * t.y > 0.f) ? ds.y * coef : 0.f;
          ds.z = ((nib & 4u) && xv.z * t.z > 0.f) ? ds.z * coef : 0.f;
          ds.w = ((nib & 8u) && xv.w * t.w > 0.f) ? ds.w * coef : 0.f;
          acc[j].x += ds.x * t.x; acc[j].y += ds.y * t.y; acc[j].z += ds.z * t.z; acc[j].w += ds.w * t.w;
          store_q<BF16>(q_buf, qrow + 128 * j, make_float4(ds.x * xv.x, ds.y * xv.y, ds.z * xv.z, ds.w * xv.w));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < CPL; ++j)
      if (EXACT || lane + 32 * j < nchunk) act_st4_stream<ABF>(gx, (size_t)u * d + 4 * (lane + 32 * j), acc[j]);
  }
}

// (2) Edge-table gradient: dT[k] = sum of the q rows whose edge has distance k. One CTA per work item (a slice
// of dist_perm inside one distance group); thread = channel chunk, rows added in slice order -> deterministic.
constexpr int kQThreads = 128;
template <bool BF16, int CPT>
__global__ void __launch_bounds__(kQThreads) dist_reduce_kernel(const void* __restrict__ q_buf,
                                                                const int* __restrict__ dist_perm,
                                                                const int4* __restrict__ items, int d,
                                                                float* __restrict__ partials) {
  const int4 it = __ldg(items + blockIdx.x);   // {dist, begin, end, 0}
  const int nchunk = d >> 2;
  float4 acc[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int kRows = 8;
  for (int r0 = it.y; r0 < it.z; r0 += kRows) {
    int pos[kRows];
#pragma unroll
    for (int i = 0; i < kRows; ++i) pos[i] = r0 + i < it.z ? __ldg(dist_perm + r0 + i) : -1;
    float4 v[kRows][CPT];
#pragma unroll
    for (int i = 0; i < kRows; ++i)
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = threadIdx.x + kQThreads * j;
        v[i][j] = (pos[i] >= 0 && c < nchunk) ? load_grad4<BF16>(q_buf, (size_t)pos[i] * d + 4 * c)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int i = 0; i < kRows; ++i)
#pragma unroll
      for (int j = 0; j < CPT; ++j) { acc[j].x += v[i][j].x; acc[j].y += v[i][j].y; acc[j].z += v[i][j].z; acc[j].w += v[i][j].w; }
  }
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const int c = threadIdx.x + kQThreads * j;
    if (c < nchunk) reinterpret_cast<float4*>(partials + (size_t)blockIdx.x * d)[c] = acc[j];
  }
}

static int check_csr(const pb_csr_t* g, int d, const char* who) {
  PB_REQUIRE(g && g->in_ptr && g->in_edge && g->out_ptr && g->out_rec, "%s: incomplete CSR plan", who);
  PB_REQUIRE(g->n_nodes > 0 && g->n_relations > 0 && g->n_relations < 32, "%s: empty graph / too many relations", who);
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

extern "C" int pb_edge_table_bwd(const float* dtable_partials, const int32_t* dist_item_ptr, int32_t d,
                                 float* g_nn_weight, float* g_nn_bias, pb_stream_t stream) {
  PB_REQUIRE(dtable_partials && dist_item_ptr && g_nn_weight && g_nn_bias && d > 0, "pb_edge_table_bwd: bad arguments");
  edge_table_bwd_kernel<<<dim3((d + 31) / 32, PB_N_DISTS), dim3(32, 8), 0, as_stream(stream)>>>(
      dtable_partials, dist_item_ptr, d, g_nn_weight);
  PB_LAUNCH_CHECK();
  edge_table_bias_kernel<<<(d + 127) / 128, 128, 0, as_stream(stream)>>>(g_nn_weight, d, g_nn_bias);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" size_t pb_dropout_bits_bytes(int64_t n_edges, int32_t d) {
  if (n_edges <= 0 || d <= 0) return 0;
  return (size_t)n_edges * keep_words(d) * sizeof(uint16_t);
}

extern "C" int pb_dropout_bits(int64_t n_edges, int32_t d, float p_drop, uint64_t seed, void* keep_bits,
                               pb_stream_t stream) {
  PB_REQUIRE(keep_bits && n_edges >= 0 && d > 0 && d % 4 == 0 && d <= 1024, "pb_dropout_bits: bad arguments");
  PB_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "pb_dropout_bits: p_drop out of range");
  if (n_edges == 0) return PB_OK;
  const int64_t total = n_edges * keep_words(d);
  const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
  uint16_t* out = reinterpret_cast<uint16_t*>(keep_bits);
  const uint32_t th = dropout_thresh(p_drop);
  cudaStream_t st = as_stream(stream);
  if (d == 512) dropout_bits_kernel<1, true><<<grid, 256, 0, st>>>(n_edges, d, th, seed, out);
  else if (d <= 512) dropout_bits_kernel<1, false><<<grid, 256, 0, st>>>(n_edges, d, th, seed, out);
  else if (d == 1024) dropout_bits_kernel<2, true><<<grid, 256, 0, st>>>(n_edges, d, th, seed, out);
  else dropout_bits_kernel<2, false><<<grid, 256, 0, st>>>(n_edges, d, th, seed, out);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

template <bool BF16, bool DROP, bool ABF>
static int launch_agg_fwd(const pb_csr_t* g, const void* x, int d, const float* table, void* a_hi, void* a_lo,
                          int64_t lda, const uint16_t* bits, float scale, cudaStream_t st) {
  const int cpl = (d + 127) / 128;
  const bool exact = d % 128 == 0 && (cpl == 1 || cpl == 2 || cpl == 4 || cpl == 8);
  const int threads = 256;
  const int64_t want = (g->n_nodes * 32 + threads - 1) / threads;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count() * 8 * 4));
#define PB_AGG_FWD(CPL, EX)                                                                                      \
  agg_fwd_kernel<BF16, DROP, CPL, EX, ABF><<<grid, threads, 0, st>>>(g->in_ptr, g->in_edge, g->in_eid, x, table, a_hi, \
                                                                a_lo, lda, g->n_nodes, d, g->n_relations, bits, scale, \
                                                                g->node_order)
  if (exact) {
    if (cpl == 1) PB_AGG_FWD(1, true);
    else if (cpl == 2) PB_AGG_FWD(2, true);
    else if (cpl == 4) PB_AGG_FWD(4, true);
    else PB_AGG_FWD(8, true);
  } else {
    if (cpl <= 1) PB_AGG_FWD(1, false);
    else if (cpl <= 2) PB_AGG_FWD(2, false);
    else if (cpl <= 4) PB_AGG_FWD(4, false);
    else PB_AGG_FWD(8, false);
  }
#undef PB_AGG_FWD
  PB_LAUNCH_CHECK();
  return PB_OK;
}

// bf16 activation storage goes with the bf16 operand mode only
static int check_act_dtype(int32_t dtype, int32_t act_dtype, const char* who) {
  PB_REQUIRE(act_dtype == PB_F32 || (act_dtype == PB_BF16 && dtype == PB_BF16),
             "%s: act_dtype %d (bf16 activations need the PB_BF16 operand mode)", who, act_dtype);
  return PB_OK;
}

extern "C" int pb_agg_fwd(const pb_csr_t* csr, const void* x, int32_t d, const float* table, void* a_hi, void* a_lo,
                          int64_t lda, int32_t dtype, const void* keep_bits, float p_drop, int32_t act_dtype,
                          pb_stream_t stream) {
  int rc = check_csr(csr, d, "pb_agg_fwd");
  if (rc) return rc;
  if ((rc = check_act_dtype(dtype, act_dtype, "pb_agg_fwd"))) return rc;
  PB_REQUIRE(x && table && a_hi, "pb_agg_fwd: null pointer");
  PB_REQUIRE(dtype == PB_BF16 || (dtype == PB_F32 && a_lo), "pb_agg_fwd: PB_F32 needs a_lo");
  PB_REQUIRE(lda >= (int64_t)(csr->n_relations + 1) * d && lda % 8 == 0, "pb_agg_fwd: bad lda");
  PB_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "pb_agg_fwd: p_drop out of range");
  PB_REQUIRE(!keep_bits || csr->in_eid, "pb_agg_fwd: dropout needs in_eid");
  PB_REQUIRE(p_drop == 0.f || keep_bits, "pb_agg_fwd: p_drop > 0 needs keep_bits (pb_dropout_bits)");
  cudaStream_t st = as_stream(stream);
  const uint16_t* bits = p_drop > 0.f ? reinterpret_cast<const uint16_t*>(keep_bits) : nullptr;
  const float scale = 1.f / (1.f - p_drop);
  if (dtype == PB_BF16 && act_dtype == PB_BF16)
    return bits ? launch_agg_fwd<true, true, true>(csr, x, d, table, a_hi, a_lo, lda, bits, scale, st)
                : launch_agg_fwd<true, false, true>(csr, x, d, table, a_hi, a_lo, lda, bits, scale, st);
  if (dtype == PB_BF16)
    return bits ? launch_agg_fwd<true, true, false>(csr, x, d, table, a_hi, a_lo, lda, bits, scale, st)
                : launch_agg_fwd<true, false, false>(csr, x, d, table, a_hi, a_lo, lda, bits, scale, st);
  return bits ? launch_agg_fwd<false, true, false>(csr, x, d, table, a_hi, a_lo, lda, bits, scale, st)
              : launch_agg_fwd<false, false, false>(csr, x, d, table, a_hi, a_lo, lda, bits, scale, st);
}

template <bool BF16, bool DROP, bool ABF>
static int launch_agg_bwd(const pb_csr_t* g, const void* x, int d, const float* table, const void* d_a, int64_t ldda,
                          const void* gy_res, void* gx, void* q_buf, float* dtp, const uint16_t* bits, float scale,
                          cudaStream_t st) {
  const int cpl = (d + 127) / 128;
  const bool exact = d % 128 == 0 && (cpl == 1 || cpl == 2 || cpl == 4 || cpl == 8);
  const int threads = 256;
  const int64_t want = (g->n_nodes * 32 + threads - 1) / threads;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count() * 8 * 4));
  const int4* recs = reinterpret_cast<const int4*>(g->out_rec);
#define PB_AGG_BWD(CPL, EX)                                                                                        \
  agg_bwd_dx_kernel<BF16, DROP, CPL, EX, ABF><<<grid, threads, 0, st>>>(g->out_ptr, recs, x, table, d_a, ldda, gy_res,   \
                                                                   gx, q_buf, g->n_nodes, d, g->n_relations, bits, scale, \
                                                                   g->node_order)
  if (exact) {
    if (cpl == 1) PB_AGG_BWD(1, true);
    else if (cpl == 2) PB_AGG_BWD(2, true);
    else if (cpl == 4) PB_AGG_BWD(4, true);
    else PB_AGG_BWD(8, true);
  } else {
    if (cpl <= 1) PB_AGG_BWD(1, false);
    else if (cpl <= 2) PB_AGG_BWD(2, false);
    else if (cpl <= 4) PB_AGG_BWD(4, false);
    else PB_AGG_BWD(8, false);
  }
#undef PB_AGG_BWD
  PB_LAUNCH_CHECK();
  const int4* items = reinterpret_cast<const int4*>(g->dist_items);
  if (d <= 4 * kQThreads)
    dist_reduce_kernel<BF16, 1><<<PB_DIST_ITEMS, kQThreads, 0, st>>>(q_buf, g->dist_perm, items, d, dtp);
  else
    dist_reduce_kernel<BF16, 2><<<PB_DIST_ITEMS, kQThreads, 0, st>>>(q_buf, g->dist_perm, items, d, dtp);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_agg_bwd(const pb_csr_t* csr, const void* x, int32_t d, const float* table, const void* d_a,
                          int64_t ldda, int32_t dtype, const void* gy_res, void* gx, void* q_buf,
                          float* dtable_partials, const void* keep_bits, float p_drop, int32_t act_dtype,
                          pb_stream_t stream) {
  int rc = check_csr(csr, d, "pb_agg_bwd");
  if (rc) return rc;
  if ((rc = check_act_dtype(dtype, act_dtype, "pb_agg_bwd"))) return rc;
  PB_REQUIRE(x && table && d_a && gx && q_buf && dtable_partials, "pb_agg_bwd: null pointer");
  PB_REQUIRE(csr->dist_perm && csr->dist_items && csr->dist_item_ptr, "pb_agg_bwd: CSR plan lacks the distance grouping");
  PB_REQUIRE(dtype == PB_BF16 || dtype == PB_F32, "pb_agg_bwd: bad dtype");
  PB_REQUIRE(ldda >= (int64_t)(csr->n_relations + 1) * d && ldda % 8 == 0, "pb_agg_bwd: bad ldda");
  PB_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "pb_agg_bwd: p_drop out of range");
  PB_REQUIRE(p_drop == 0.f || keep_bits, "pb_agg_bwd: p_drop > 0 needs keep_bits (pb_dropout_bits)");
  cudaStream_t st = as_stream(stream);
  const uint16_t* bits = p_drop > 0.f ? reinterpret_cast<const uint16_t*>(keep_bits) : nullptr;
  const float scale = 1.f / (1.f - p_drop);
#define PB_AGG_BWD_CALL(BF, DR, AB) \
  launch_agg_bwd<BF, DR, AB>(csr, x, d, table, d_a, ldda, gy_res, gx, q_buf, dtable_partials, bits, scale, st)
  if (dtype == PB_BF16 && act_dtype == PB_BF16) return bits ? PB_AGG_BWD_CALL(true, true, true) : PB_AGG_BWD_CALL(true, false, true);
  if (dtype == PB_BF16) return bits ? PB_AGG_BWD_CALL(true, true, false) : PB_AGG_BWD_CALL(true, false, false);
  return bits ? PB_AGG_BWD_CALL(false, true, false) : PB_AGG_BWD_CALL(false, false, false);
#undef PB_AGG_BWD_CALL
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
