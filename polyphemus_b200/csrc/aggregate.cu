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
#include <stdlib.h>

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

// Pipelined forward for d = 128 * CPL <= 512 on plans with visit_meta. The kernel above walks, per node, a chain of
// dependent loads (node_order -> in_ptr -> in_edge -> rows: four L2 round trips of ~0.7 us with one row in flight per
// warp), which is what bounds it (28 KB in flight per SM ~ 4 TB/s). Here a persistent warp keeps the chain of the next
// two nodes in flight (metadata two nodes ahead, segment offsets + edge records one node ahead) and the rows of the
// next in-edge in registers while it consumes the current one. GCL.message's dropout is applied to the raw feature
// words (AND with a mask from a 16-entry shared-memory table: 5 instructions per 4 channels instead of 16); a dropped
// channel is an exact +0 from there on. (Handing the nodes out dynamically, in chunks through an atomic counter, so that
// a late CTA would not add a second wave, measured 163 us against 157 for this static interleaving and no gain inside
// the training step; the launch bound keeps PB_FWD_PIPE_CTAS CTAs per SM resident, which the grid size assumes.)
#ifndef PB_FWD_PIPE_CTAS
#define PB_FWD_PIPE_CTAS 3      // resident CTAs per SM the persistent grid is sized for (79 registers at d = 512)
#endif
template <bool ABF> struct XRaw { using type = float4; };
template <> struct XRaw<true> { using type = uint2; };

template <bool BF16, bool DROPOUT, int CPL, bool ABF>
__global__ void __launch_bounds__(256, PB_FWD_PIPE_CTAS) agg_fwd_pipe_kernel(const int4* __restrict__ visit_meta, const int* __restrict__ in_ptr,
                                                           const int* __restrict__ in_edge, const int* __restrict__ in_eid,
                                                           const void* __restrict__ x, const float* __restrict__ table,
                                                           void* __restrict__ a_hi, void* __restrict__ a_lo, int64_t lda,
                                                           int64_t n_nodes, int n_rel, int n_edges,
                                                           const uint16_t* __restrict__ keep_bits, float keep_scale) {
  constexpr uint32_t kFull = 0xffffffffu;
  constexpr int d = 128 * CPL;
  static_assert(CPL <= 4, "one keep-bit word per lane");
  using Raw = typename XRaw<ABF>::type;
  __shared__ uint4 lut[16];                                  // 4 keep bits -> AND masks (bf16: .x/.y pairs, fp32: 4 words)
  if (threadIdx.x < 16) {
    const uint32_t b = threadIdx.x;
    if constexpr (ABF)
      lut[b] = make_uint4(((b & 1u) ? 0xFFFFu : 0u) | ((b & 2u) ? 0xFFFF0000u : 0u),
                          ((b & 4u) ? 0xFFFFu : 0u) | ((b & 8u) ? 0xFFFF0000u : 0u), 0u, 0u);
    else
      lut[b] = make_uint4((b & 1u) ? ~0u : 0u, (b & 2u) ? ~0u : 0u, (b & 4u) ? ~0u : 0u, (b & 8u) ? ~0u : 0u);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  struct Recs { int ptr; uint32_t pk, eid; };
  auto load_meta = [&](int64_t vi) { return vi < n_nodes ? __ldg(visit_meta + vi) : make_int4(0, 0, 0, 0); };
  // segment offsets of the node (lanes 0..n_rel) and its first 32 in-edge records (reading past the node's last edge
  // is harmless: the index is clamped to the array, the values are never used)
  auto load_recs = [&](const int4& m) {
    Recs r;
    r.ptr = lane <= n_rel ? __ldg(in_ptr + (int64_t)m.x * n_rel + lane) : 0;
    const int i = min(m.w + lane, n_edges - 1);
    r.pk = (uint32_t)__ldg(in_edge + i);
    r.eid = 0;
    if constexpr (DROPOUT) r.eid = (uint32_t)__ldg(in_eid + i);
    return r;
  };
  auto ld_row = [&](Raw (&raw)[CPL], size_t elem) {
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      if constexpr (ABF) raw[j] = __ldg(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(x) + elem + 128 * j));
      else raw[j] = ldg4(static_cast<const float*>(x) + elem + 128 * j);
    }
  };
  auto unpack = [&](const Raw& r) {
    if constexpr (ABF) {
      const float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y);
      return make_float4(a.x, a.y, b.x, b.y);
    } else {
      return r;
    }
  };
  int4 m_cur = load_meta(w0), m_nxt = load_meta(w0 + n_warps), m_nn;
  Recs rc = load_recs(m_cur), rn;
  for (int64_t vi = w0; vi < n_nodes; vi += n_warps) {
    m_nn = load_meta(vi + 2 * n_warps);
    rn = load_recs(m_nxt);
    const int64_t v = m_cur.x;
    const size_t row = (size_t)v * lda;
    const int beg_all = m_cur.w;
    const int end_all = __shfl_sync(kFull, rc.ptr, n_rel);
    int base = beg_all;
    uint32_t my_pk = rc.pk, my_eid = rc.eid;
    Raw ra[CPL], rb[CPL];
    uint32_t pka = 0, pkb = 0, kwa = 0, kwb = 0;
    auto fetch = [&](Raw (&raw)[CPL], uint32_t& pk, uint32_t& kw, int e) {
      if (e - base == 32) {  // warp-uniform: next batch of records (in-degree > 32 only)
        base = e;
        const int i = min(base + lane, n_edges - 1);
        my_pk = (uint32_t)__ldg(in_edge + i);
        if constexpr (DROPOUT) my_eid = (uint32_t)__ldg(in_eid + i);
      }
      pk = __shfl_sync(kFull, my_pk, e - base);
      ld_row(raw, (size_t)(pk & 0x03FFFFFFu) * d + 4 * lane);
      if constexpr (DROPOUT) {
        const uint32_t eid = __shfl_sync(kFull, my_eid, e - base);
        kw = __ldg(keep_bits + (size_t)eid * 32 + lane);
      }
    };
    if (beg_all < end_all) fetch(ra, pka, kwa, beg_all);
    // root block: the node's own features, converted to the operand dtype
    {
      Raw rr[CPL];
      ld_row(rr, (size_t)v * d + 4 * lane);
#pragma unroll
      for (int j = 0; j < CPL; ++j)
        store_operand<BF16>(a_hi, a_lo, row + (size_t)n_rel * d + 4 * (lane + 32 * j), unpack(rr[j]));
    }
    int e = beg_all;
    for (int r = 0; r < n_rel; ++r) {
      const int seg_end = __shfl_sync(kFull, rc.ptr, r + 1);
      const int cnt = seg_end - e;
      float4 acc[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (; e < seg_end; ++e) {
        if (e + 1 < end_all) fetch(rb, pkb, kwb, e + 1);           // next in-edge (of any segment) in flight
        const float* trow = table + (size_t)(pka >> 26) * d + 4 * lane;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          Raw xr = ra[j];
          if constexpr (DROPOUT) {
            const uint4 mk = lut[(kwa >> (4 * j)) & 15u];
            if constexpr (ABF) { xr.x &= mk.x; xr.y &= mk.y; }
            else {
              xr.x = __uint_as_float(__float_as_uint(xr.x) & mk.x); xr.y = __uint_as_float(__float_as_uint(xr.y) & mk.y);
              xr.z = __uint_as_float(__float_as_uint(xr.z) & mk.z); xr.w = __uint_as_float(__float_as_uint(xr.w) & mk.w);
            }
          }
          const float4 xs = unpack(xr);
          const float4 t = ldg4(trow + 128 * j);
          const float4 m = make_float4(fmaxf(xs.x * t.x, 0.f), fmaxf(xs.y * t.y, 0.f), fmaxf(xs.z * t.z, 0.f),
                                       fmaxf(xs.w * t.w, 0.f));
          if constexpr (DROPOUT) {
            acc[j].x = fmaf(m.x, keep_scale, acc[j].x); acc[j].y = fmaf(m.y, keep_scale, acc[j].y);
            acc[j].z = fmaf(m.z, keep_scale, acc[j].z); acc[j].w = fmaf(m.w, keep_scale, acc[j].w);
          } else {
            acc[j].x += m.x; acc[j].y += m.y; acc[j].z += m.z; acc[j].w += m.w;
          }
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j) ra[j] = rb[j];
        pka = pkb; kwa = kwb;
      }
      const float inv = cnt > 1 ? 1.0f / (float)cnt : 1.0f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const float4 h = make_float4(acc[j].x * inv, acc[j].y * inv, acc[j].z * inv, acc[j].w * inv);
        store_operand<BF16>(a_hi, a_lo, row + (size_t)r * d + 4 * (lane + 32 * j), h);
      }
    }
    m_cur = m_nxt; m_nxt = m_nn; rc = rn;
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

// ---------------------------------------------------------------------------------------------- fused backward
// Scatter-by-source AND the edge-table gradient in one pass, no E x d intermediate (the legacy path above writes the
// rows q_e = ds_e * x[src_e] to HBM and reads them back grouped by distance: 2 x E x d bytes of pure overhead).
//
// The obstacle is the 32-bin reduction dT[dist_e] += q_e. Here a CTA owns a contiguous range of source nodes (in
// visiting order) and its threads split the CHANNELS: thread t owns channels [t*CH, t*CH + CH) of every row the CTA
// touches — of the gradient rows it gathers, of gx, and of a [32, d] fp32 accumulator dT in shared memory. Nobody
// else ever touches those columns, so the accumulation needs no atomics, no barriers and no inter-warp ordering: the
// warps of a CTA run completely independently (each is a channel slice of the same source stream), and the order of
// additions into every accumulator word is the order of the edges — bit-reproducible. Consecutive edges with the
// same distance (the ONSET edges of a source all have distance 0, its NEXT edges share one distance) are summed in
// registers first. Every CTA leaves its [32, d] partial in `partials`; pb_edge_table_bwd_fused adds them in fixed
// order.
//
// Latency hiding: a warp sees only d / n_warps channels of a row (256 B of a 1 KB bf16 row), so it must keep many
// rows in flight. The work of a source range is therefore laid out by the plan as ONE flat record stream
// (pb_csr_bwd_stream): per source {x row, root-block row, residual row, out-edge rows...}, one 16-byte record each.
// A warp loads 32 records per coalesced load (two chunks ahead) and keeps a ring of kRing rows in flight in
// registers: slot s of the ring is consumed (record p) and immediately refilled with the row of record p + kRing.
// The ring is indexed statically (the trip over its slots is fully unrolled), so nothing is spilled or rotated.
constexpr int kRing = 16;

template <int W> struct RawW { uint32_t w[W]; };
template <int W>
__device__ __forceinline__ RawW<W> ld_raw(const void* p) {   // read-once data: keep it out of L1 (the edge table lives there)
  RawW<W> r;
  if constexpr (W == 1) {
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r.w[0]) : "l"(p));
  } else if constexpr (W == 2) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];" : "=r"(r.w[0]), "=r"(r.w[1]) : "l"(p));
  } else {
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3])
                 : "l"(p));
  }
  return r;
}
// CH values stored as bf16 (BF) or fp32 in the first CH * (BF ? 2 : 4) / 4 words
template <bool BF, int CH, int W>
__device__ __forceinline__ void unpack_raw(const RawW<W>& r, float v[CH]) {
  if constexpr (BF) {
#pragma unroll
    for (int i = 0; i < CH / 2; ++i) { const float2 f = unpack_bf16x2(r.w[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  } else {
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(r.w[i]);
  }
}
template <bool BF, int CH, int W>
__device__ __forceinline__ RawW<W> ld_vals(const void* base, size_t elem) {
  constexpr int N = CH * (BF ? 2 : 4) / 4;
  const RawW<N> t = ld_raw<N>(static_cast<const char*>(base) + elem * (BF ? 2 : 4));
  RawW<W> r;
#pragma unroll
  for (int i = 0; i < W; ++i) r.w[i] = i < N ? t.w[i < N ? i : 0] : 0u;
  return r;
}
template <bool BF, int CH>
__device__ __forceinline__ void store_vals(void* base, size_t elem, const float v[CH]) {
  if constexpr (BF) {
    if constexpr (CH == 4) st_stream2(static_cast<__nv_bfloat16*>(base) + elem, make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3])));
    else {
      const uint32_t u = pack_bf16x2(v[0], v[1]);
      asm volatile("st.global.L1::no_allocate.b32 [%0], %1;" ::"l"(static_cast<__nv_bfloat16*>(base) + elem), "r"(u) : "memory");
    }
  } else {
    if constexpr (CH == 4) st_stream4(static_cast<float*>(base) + elem, make_float4(v[0], v[1], v[2], v[3]));
    else asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(static_cast<float*>(base) + elem), "f"(v[0]), "f"(v[1]) : "memory");
  }
}

template <bool BF16, bool DROPOUT, int CH, bool ABF>
__global__ void __launch_bounds__(256) agg_bwd_fused_kernel(
    const int4* __restrict__ stream, const int* __restrict__ visit_edge_ptr, const void* __restrict__ x,
    const float* __restrict__ table, const void* __restrict__ d_a, const void* __restrict__ gy_res,
    void* __restrict__ gx, float* __restrict__ partials, int64_t n_nodes, int d,
    const uint16_t* __restrict__ keep_bits, float keep_scale) {
  constexpr uint32_t kFull = 0xffffffffu;
  constexpr int GW = CH * (BF16 ? 2 : 4) / 4, XW = CH * (ABF ? 2 : 4) / 4, W = GW > XW ? GW : XW;
  extern __shared__ float dT_s[];                        // [32][d]; column c belongs to thread c / CH
  const int lane = threadIdx.x & 31;
  const int c0 = threadIdx.x * CH;                       // blockDim.x * CH == d
  const int g16 = (d + 511) / 512;                       // keep-bit words per edge / 32
  const int cidx = c0 >> 2;                              // 4-channel chunk of the keep-bit layout
  const int kb_word = (cidx >> 7) * 32 + (cidx & 31), kb_shift = 4 * ((cidx >> 5) & 3) + (CH == 2 ? (c0 & 2) : 0);
  float* my_t = dT_s + c0;
#pragma unroll 4
  for (int k = 0; k < PB_N_DISTS; ++k)
#pragma unroll
    for (int j = 0; j < CH; ++j) my_t[k * d + j] = 0.f;

  const int64_t per = (n_nodes + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per;
  const int64_t r1 = r0 + per < n_nodes ? r0 + per : n_nodes;

  if (r0 < r1) {
    const int64_t p0 = 3 * r0 + __ldg(visit_edge_ptr + r0), p1 = 3 * r1 + __ldg(visit_edge_ptr + r1);
    auto load_recs = [&](int64_t base) {
      return base + lane < p1 ? __ldg(stream + base + lane) : make_int4(0, kRecNop, 0, 0);
    };
    // state of the source being consumed
    float xu[CH], acc[CH], run_q[CH];
    int run_dist = 0, cur_u = 0;
#pragma unroll
    for (int j = 0; j < CH; ++j) { xu[j] = 0.f; acc[j] = 0.f; run_q[j] = 0.f; }
    auto flush_run = [&]() {
      float* p = my_t + run_dist * d;
      if constexpr (CH == 4) {
        float4 v = *reinterpret_cast<float4*>(p);
        v.x += run_q[0]; v.y += run_q[1]; v.z += run_q[2]; v.w += run_q[3];
        *reinterpret_cast<float4*>(p) = v;
      } else {
        float2 v = *reinterpret_cast<float2*>(p);
        v.x += run_q[0]; v.y += run_q[1];
        *reinterpret_cast<float2*>(p) = v;
      }
#pragma unroll
      for (int j = 0; j < CH; ++j) run_q[j] = 0.f;
    };
    RawW<W> ring[kRing];
    uint32_t ring_kb[kRing];
    // issue the row of record (rx, ry, rz) into ring slot s
    auto issue = [&](RawW<W>& slot, uint32_t& kb, int rx, int ry, int rz) {
      const int kind = ry & 7;
      const size_t off = (size_t)(uint32_t)rx * d + c0;
      if (kind == kRecEdge || kind == kRecRoot) {                 // warp-uniform
        slot = ld_vals<BF16, CH, W>(d_a, off);
        if constexpr (DROPOUT)
          if (kind == kRecEdge) kb = __ldg(keep_bits + (size_t)(uint32_t)rz * g16 * 32 + kb_word);
      } else if (kind == kRecX) {
        slot = ld_vals<ABF, CH, W>(x, off);
      } else if (kind == kRecRes) {
        if (gy_res) slot = ld_vals<ABF, CH, W>(gy_res, off);
      }
    };
    auto consume = [&](const RawW<W>& slot, uint32_t kb, int ry, int rz, int rw) {
      const int kind = ry & 7;
      float v[CH];
      if (kind == kRecEdge) {
        const int dist = (ry >> 8) & (PB_N_DISTS - 1);
        float t[CH];
        if constexpr (CH == 4) { const float4 tv = ldg4(table + (size_t)dist * d + c0); t[0] = tv.x; t[1] = tv.y; t[2] = tv.z; t[3] = tv.w; }
        else { const float2 tv = __ldg(reinterpret_cast<const float2*>(table + (size_t)dist * d + c0)); t[0] = tv.x; t[1] = tv.y; }
        unpack_raw<BF16, CH, W>(slot, v);
        float coef = __int_as_float(rw);                          // 1 / |segment|, from the plan
        if constexpr (DROPOUT) coef *= keep_scale;
        if (dist != run_dist) { flush_run(); run_dist = dist; }   // warp-uniform
        const uint32_t bits = DROPOUT ? (kb >> kb_shift) : 0xFu;
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const bool keep = ((bits >> j) & 1u) && xu[j] * t[j] > 0.f;
          const float ds = keep ? v[j] * coef : 0.f;
          acc[j] += ds * t[j];
          run_q[j] += ds * xu[j];
        }
      } else if (kind == kRecX) {
        unpack_raw<ABF, CH, W>(slot, xu);
        cur_u = rz;                                               // header records carry the source row in z
      } else if (kind == kRecRoot) {
        unpack_raw<BF16, CH, W>(slot, acc);
      } else if (kind == kRecRes) {
        if (gy_res) {
          unpack_raw<ABF, CH, W>(slot, v);
#pragma unroll
          for (int j = 0; j < CH; ++j) acc[j] += v[j];
        }
      }
      if (ry & kRecLast) store_vals<ABF, CH>(gx, (size_t)(uint32_t)cur_u * d + c0, acc);
    };

    int4 recs_cur = load_recs(p0), recs_nxt = load_recs(p0 + 32);
#pragma unroll
    for (int s = 0; s < kRing; ++s)
      issue(ring[s], ring_kb[s], __shfl_sync(kFull, recs_cur.x, s), __shfl_sync(kFull, recs_cur.y, s),
            __shfl_sync(kFull, recs_cur.z, s));
    for (int64_t base = p0; base < p1; base += 32) {
      const int4 recs_nn = load_recs(base + 64);
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        // records base + 16 half + s are in the ring; refill with records base + 16 half + 16 + s
        const int4 nx = half ? recs_nxt : recs_cur;
        const int lc = 16 * half, ln = 16 - 16 * half;
#pragma unroll
        for (int s = 0; s < kRing; ++s) {
          consume(ring[s], ring_kb[s], __shfl_sync(kFull, recs_cur.y, lc + s), __shfl_sync(kFull, recs_cur.z, lc + s),
                  __shfl_sync(kFull, recs_cur.w, lc + s));
          issue(ring[s], ring_kb[s], __shfl_sync(kFull, nx.x, ln + s), __shfl_sync(kFull, nx.y, ln + s),
                __shfl_sync(kFull, nx.z, ln + s));
        }
      }
      recs_cur = recs_nxt;
      recs_nxt = recs_nn;
    }
    flush_run();
  }
  __syncwarp();
  float* out = partials + (size_t)blockIdx.x * PB_N_DISTS * d + c0;
#pragma unroll 4
  for (int k = 0; k < PB_N_DISTS; ++k)
#pragma unroll
    for (int j = 0; j < CH; ++j) out[(size_t)k * d + j] = my_t[k * d + j];
}

// The record stream of the fused backward, in visiting order: per source {X, ROOT, RES, out-edges...}.
//   x = row-block index: the row u for X / RES (rows of x / gy_res), u * (R+1) + R for ROOT and dst * (R+1) + slot for
//       an edge (blocks of d elements inside d_a [n, (R+1) d])
//   y = kind | kRecLast on the source's last record | dist << 8      z = edge id (edge) / row u (header records)
//   w = 1 / |segment(dst, slot)| as float bits (edge)
__global__ void bwd_stream_kernel(const int4* __restrict__ visit_meta, const int* __restrict__ visit_edge_ptr,
                                  const int4* __restrict__ out_rec, int n_rel, int64_t n_nodes, int4* __restrict__ stream) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const int4 m = visit_meta[i];                         // {row, first out-edge, out-degree, -}
  int4* o = stream + 3 * i + visit_edge_ptr[i];
  const int u = m.x, deg = m.z;
  o[0] = make_int4(u, kRecX, u, 0);
  o[1] = make_int4(u * (n_rel + 1) + n_rel, kRecRoot, u, 0);
  o[2] = make_int4(u, kRecRes | (deg == 0 ? kRecLast : 0), u, 0);
  for (int s = 0; s < deg; ++s) {
    const int4 r = out_rec[m.y + s];                     // {dst, slot | dist << 8, eid, |segment|}
    const float coef = r.w > 1 ? 1.0f / (float)r.w : 1.0f;
    o[3 + s] = make_int4(r.x * (n_rel + 1) + (r.y & 0xff), kRecEdge | (s == deg - 1 ? kRecLast : 0) | (r.y & 0x1f00), r.z,
                         __float_as_int(coef));
  }
}

// partials [P][32][d] (one block per CTA of agg_bwd_fused_kernel) -> g_w [d][32]: grid (d/32, 32 distances), block =
// 32 channels x 8 lanes over P, lanes combined in fixed order.
__global__ void __launch_bounds__(256) edge_table_bwd_fused_kernel(const float* __restrict__ partials, int n_part, int d,
                                                                   float* __restrict__ g_w) {
  __shared__ float red[8][33];
  const int ci = threadIdx.x, j = threadIdx.y, k = blockIdx.y;
  const int c = blockIdx.x * 32 + ci;
  float s = 0.f;
  if (c < d)
    for (int p = j; p < n_part; p += 8) s += partials[((size_t)p * PB_N_DISTS + k) * d + c];
  red[j][ci] = s;
  __syncthreads();
  if (j == 0 && c < d) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][ci];
    g_w[(size_t)c * PB_N_DISTS + k] = t;
  }
}

// {row, first out-edge, out-degree, first in-edge} per visited node, in visiting order
__global__ void visit_meta_kernel(const int* __restrict__ node_order, const int* __restrict__ out_ptr,
                                  const int* __restrict__ in_ptr, int n_rel, int64_t n_nodes, int4* __restrict__ meta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const int u = node_order ? node_order[i] : (int)i;
  const int beg = out_ptr[u];
  meta[i] = make_int4(u, beg, out_ptr[u + 1] - beg, in_ptr[(size_t)u * n_rel]);
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

// PB200_AGG_FWD_PIPE=0 keeps the unpipelined forward (A/B measurements)
static bool fwd_pipe_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PB200_AGG_FWD_PIPE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <bool BF16, bool DROP, bool ABF>
static int launch_agg_fwd(const pb_csr_t* g, const void* x, int d, const float* table, void* a_hi, void* a_lo,
                          int64_t lda, const uint16_t* bits, float scale, cudaStream_t st) {
  const int cpl = (d + 127) / 128;
  const bool exact = d % 128 == 0 && (cpl == 1 || cpl == 2 || cpl == 4 || cpl == 8);
  const int threads = 256;
  // measured: at d = 512 with bf16 rows 157 us against 172 (LMD16 batch 256), 0.68 / 0.61 of the HBM roofline against
  // 0.63 / 0.60 at E = 1e6 / 1e7; at d = 256 (512-byte rows: the per-node work dominates) and with fp32 rows (twice the
  // registers per row in flight) the plain kernel's higher occupancy wins
  if (ABF && exact && cpl == 4 && g->visit_meta && g->n_edges > 0 && fwd_pipe_enabled()) {
    const int64_t want_p = (g->n_nodes * 32 + threads - 1) / threads;
    const unsigned grid_p = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want_p, (int64_t)sm_count() * PB_FWD_PIPE_CTAS));
#define PB_AGG_PIPE(CPL)                                                                                            \
  agg_fwd_pipe_kernel<BF16, DROP, CPL, ABF><<<grid_p, threads, 0, st>>>(                                             \
      reinterpret_cast<const int4*>(g->visit_meta), g->in_ptr, g->in_edge, g->in_eid, x, table, a_hi, a_lo, lda, g->n_nodes, \
      g->n_relations, (int)g->n_edges, bits, scale)
    PB_AGG_PIPE(4);
#undef PB_AGG_PIPE
    PB_LAUNCH_CHECK();
    return PB_OK;
  }
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

// ---- fused backward: launch geometry shared by the kernel launch and the partials-size query
static int fused_ch(int d) { return d % 128 == 0 ? 4 : 2; }
static int fused_ctas(int64_t n_nodes, int d) {
  const size_t smem = (size_t)PB_N_DISTS * d * sizeof(float);
  int per_sm = (int)((size_t)227 * 1024 / (smem + 1024));
  per_sm = std::max(1, std::min(per_sm, 8));
  const int64_t want = (n_nodes + 63) / 64;                      // at least 64 sources per CTA
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count() * per_sm));
}

// PB200_AGG_BWD_TC=0 keeps the CUDA-core variant for eligible shapes too (A/B measurements)
static bool tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PB200_AGG_BWD_TC");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

extern "C" int32_t pb_agg_bwd_num_partials(int64_t n_nodes, int32_t d, int32_t dtype) {
  if (n_nodes <= 0 || d <= 0) return 0;
  if (tc_enabled() && agg_bwd_tc_eligible(d, dtype)) return agg_bwd_tc_ctas(n_nodes);
  return fused_ctas(n_nodes, d);
}

template <bool BF16, bool DROP, int CH, bool ABF>
static int launch_agg_bwd_fused(const pb_csr_t* g, const void* x, int d, const float* table, const void* d_a,
                                const void* gy_res, void* gx, float* partials, const uint16_t* bits, float scale,
                                cudaStream_t st) {
  const size_t smem = (size_t)PB_N_DISTS * d * sizeof(float);
  static bool attr_set[64] = {};
  int dev = 0;
  PB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    PB_CUDA(cudaFuncSetAttribute(agg_bwd_fused_kernel<BF16, DROP, CH, ABF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 PB_N_DISTS * 1024 * (int)sizeof(float)));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int grid = fused_ctas(g->n_nodes, d);
  agg_bwd_fused_kernel<BF16, DROP, CH, ABF><<<grid, d / CH, smem, st>>>(
      reinterpret_cast<const int4*>(g->bwd_stream), g->visit_edge_ptr, x, table, d_a, gy_res, gx, partials, g->n_nodes, d,
      bits, scale);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_agg_bwd_fused(const pb_csr_t* csr, const void* x, int32_t d, const float* table, const void* d_a,
                                int64_t ldda, int32_t dtype, const void* gy_res, void* gx, float* dtable_partials,
                                const void* keep_bits, float p_drop, int32_t act_dtype, pb_stream_t stream) {
  int rc = check_csr(csr, d, "pb_agg_bwd_fused");
  if (rc) return rc;
  if ((rc = check_act_dtype(dtype, act_dtype, "pb_agg_bwd_fused"))) return rc;
  PB_REQUIRE(x && table && d_a && gx && dtable_partials, "pb_agg_bwd_fused: null pointer");
  PB_REQUIRE(csr->bwd_stream && csr->visit_edge_ptr, "pb_agg_bwd_fused: CSR plan lacks the record stream (pb_csr_bwd_stream)");
  PB_REQUIRE(ldda == (int64_t)(csr->n_relations + 1) * d, "pb_agg_bwd_fused: d_a must be packed (ldda == (R+1) d)");
  PB_REQUIRE((int64_t)csr->n_nodes * (csr->n_relations + 1) < ((int64_t)1 << 32), "pb_agg_bwd_fused: too many row blocks");
  PB_REQUIRE(dtype == PB_BF16 || dtype == PB_F32, "pb_agg_bwd_fused: bad dtype");
  PB_REQUIRE(ldda >= (int64_t)(csr->n_relations + 1) * d && ldda % 8 == 0, "pb_agg_bwd_fused: bad ldda");
  PB_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "pb_agg_bwd_fused: p_drop out of range");
  PB_REQUIRE(p_drop == 0.f || keep_bits, "pb_agg_bwd_fused: p_drop > 0 needs keep_bits (pb_dropout_bits)");
  cudaStream_t st = as_stream(stream);
  const uint16_t* bits = p_drop > 0.f ? reinterpret_cast<const uint16_t*>(keep_bits) : nullptr;
  const float scale = 1.f / (1.f - p_drop);
  if (tc_enabled() && agg_bwd_tc_eligible(d, dtype)) {
    PB_REQUIRE(csr->visit_meta, "pb_agg_bwd_fused: CSR plan lacks visit_meta (pb_csr_visit_meta)");
    return agg_bwd_tc_launch(csr, x, d, table, d_a, ldda, gy_res, gx, dtable_partials, bits, scale, act_dtype == PB_BF16, st);
  }
#define PB_FUSED(BF, DR, CH, AB) \
  launch_agg_bwd_fused<BF, DR, CH, AB>(csr, x, d, table, d_a, gy_res, gx, dtable_partials, bits, scale, st)
#define PB_FUSED_CH(BF, DR, AB) (fused_ch(d) == 4 ? PB_FUSED(BF, DR, 4, AB) : PB_FUSED(BF, DR, 2, AB))
  if (dtype == PB_BF16 && act_dtype == PB_BF16) return bits ? PB_FUSED_CH(true, true, true) : PB_FUSED_CH(true, false, true);
  if (dtype == PB_BF16) return bits ? PB_FUSED_CH(true, true, false) : PB_FUSED_CH(true, false, false);
  return bits ? PB_FUSED_CH(false, true, false) : PB_FUSED_CH(false, false, false);
#undef PB_FUSED_CH
#undef PB_FUSED
}

extern "C" int pb_edge_table_bwd_fused(const float* dtable_partials, int32_t n_partials, int32_t d, float* g_nn_weight,
                                       float* g_nn_bias, pb_stream_t stream) {
  PB_REQUIRE(dtable_partials && g_nn_weight && g_nn_bias && d > 0 && n_partials > 0, "pb_edge_table_bwd_fused: bad arguments");
  edge_table_bwd_fused_kernel<<<dim3((d + 31) / 32, PB_N_DISTS), dim3(32, 8), 0, as_stream(stream)>>>(
      dtable_partials, n_partials, d, g_nn_weight);
  PB_LAUNCH_CHECK();
  edge_table_bias_kernel<<<(d + 127) / 128, 128, 0, as_stream(stream)>>>(g_nn_weight, d, g_nn_bias);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_csr_bwd_stream(const pb_csr_t* csr, const int32_t* visit_edge_ptr, void* bwd_stream, pb_stream_t stream) {
  PB_REQUIRE(csr && csr->visit_meta && csr->out_rec && visit_edge_ptr && bwd_stream && csr->n_nodes > 0,
             "pb_csr_bwd_stream: bad arguments (needs visit_meta)");
  PB_REQUIRE((reinterpret_cast<uintptr_t>(bwd_stream) & 15) == 0, "pb_csr_bwd_stream: stream must be 16B aligned");
  const int64_t n = csr->n_nodes;
  bwd_stream_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const int4*>(csr->visit_meta), visit_edge_ptr, reinterpret_cast<const int4*>(csr->out_rec),
      csr->n_relations, n, reinterpret_cast<int4*>(bwd_stream));
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_csr_visit_meta(const pb_csr_t* csr, void* visit_meta, pb_stream_t stream) {
  PB_REQUIRE(csr && csr->in_ptr && csr->out_ptr && visit_meta && csr->n_nodes > 0, "pb_csr_visit_meta: bad arguments");
  PB_REQUIRE((reinterpret_cast<uintptr_t>(visit_meta) & 15) == 0, "pb_csr_visit_meta: visit_meta must be 16B aligned");
  const int64_t n = csr->n_nodes;
  visit_meta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
      csr->node_order, csr->out_ptr, csr->in_ptr, csr->n_relations, n, reinterpret_cast<int4*>(visit_meta));
  PB_LAUNCH_CHECK();
  return PB_OK;
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
