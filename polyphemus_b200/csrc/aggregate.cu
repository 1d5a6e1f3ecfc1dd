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

// ---------------------------------------------------------------------------------------------- dropout bits
// GCL.message's dropout (model.py:133). The keep decision of every (edge, channel) is drawn once per layer call
// by a counter-based hash (common.cuh) and packed 1 bit per channel; forward, operand recompute and backward all
// read the same bits, so the hot kernels carry no RNG arithmetic.
// Layout: u16 word [(e * G16 + jj) * 32 + l] holds chunks c = l + 32 * (4 jj + nib), nib = 0..3, 4 bits each
// (chunk = 4 consecutive channels) — i.e. exactly the channels lane l of a warp owns in agg_fwd.
static inline int keep_words(int d) { return ((d + 511) / 512) * 32; }   // u16 words per edge

__global__ void dropout_bits_kernel(int64_t n_edges, int d, uint32_t thresh16, uint64_t seed, uint16_t* __restrict__ bits) {
  const int nchunk = d >> 2;
  const int g16 = (d + 511) / 512;
  const int64_t total = n_edges * g16 * 32;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(i & 31);
    const int64_t q = i >> 5;
    const int jj = (int)(q % g16);
    const int64_t e = q / g16;
    uint32_t w = 0;
#pragma unroll
    for (int nib = 0; nib < 4; ++nib) {
      const int c = l + 32 * (4 * jj + nib);
      if (c < nchunk) {
        bool k[4];
        dropout_keep4(seed, (uint32_t)e, (uint32_t)c, thresh16, k);
        w |= (uint32_t)(k[0] | (k[1] << 1) | (k[2] << 2) | (k[3] << 3)) << (4 * nib);
      }
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
template <bool BF16, bool DROPOUT, int CPL, bool EXACT>
__global__ void __launch_bounds__(256) agg_fwd_kernel(const int* __restrict__ in_ptr, const int* __restrict__ in_edge,
                                                      const int* __restrict__ in_eid, const float* __restrict__ x,
                                                      const float* __restrict__ table, void* __restrict__ a_hi,
                                                      void* __restrict__ a_lo, int64_t lda, int64_t n_nodes, int d,
                                                      int n_rel, const uint16_t* __restrict__ keep_bits, float keep_scale) {
  constexpr uint32_t kFull = 0xffffffffu;
  constexpr int G16 = (CPL + 3) / 4;
  const int lane = threadIdx.x & 31;
  const int nchunk = d >> 2;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < n_nodes; v += n_warps) {
    const size_t row = (size_t)v * lda;
    const int my_ptr = lane <= n_rel ? __ldg(in_ptr + v * n_rel + lane) : 0;
    const int beg_all = __shfl_sync(kFull, my_ptr, 0);
    const int end_all = __shfl_sync(kFull, my_ptr, n_rel);
    int base = beg_all;
    uint32_t my_pk = base + lane < end_all ? (uint32_t)__ldg(in_edge + base + lane) : 0u;
    uint32_t my_eid = 0;
    if constexpr (DROPOUT) my_eid = base + lane < end_all ? (uint32_t)__ldg(in_eid + base + lane) : 0u;
    // root block: the node's own features, converted to the operand dtype
    const float* xrow = x + (size_t)v * d + 4 * lane;
#pragma unroll
    for (int j = 0; j < CPL; ++j)
      if (EXACT || lane + 32 * j < nchunk)
        store_operand<BF16>(a_hi, a_lo, row + (size_t)n_rel * d + 4 * (lane + 32 * j), ldg4(xrow + 128 * j));
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
        const float* srow = x + (size_t)(pk & 0x03FFFFFFu) * d + 4 * lane;
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
            const float4 xs = ldg4(srow + 128 * j);
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
// Software pipeline: while node u is processed, the rows and edge records of the CTA's next node are in
// flight; inside a node the gathers of up to kBatch edges are issued before any of them is consumed.
template <bool BF16, bool DROPOUT, int CPT>
__global__ void __launch_bounds__(kBwdThreads) agg_bwd_kernel(
    const int* __restrict__ out_ptr, const int4* __restrict__ out_rec, const float* __restrict__ x,
    const float* __restrict__ table, const void* __restrict__ d_a, int64_t ldda, const float* __restrict__ gy_res,
    float* __restrict__ gx, float* __restrict__ dt_partials, int64_t n_nodes, int d, int n_rel, int tpn,
    const uint16_t* __restrict__ keep_bits, float keep_scale) {
  extern __shared__ float dts[];  // [G][32][d]
  const int groups = kBwdThreads / tpn;
  const int g = threadIdx.x / tpn, tc = threadIdx.x % tpn;
  const int sw = tpn < 32 ? tpn : 32;          // shuffle width: threads of one node inside a warp
  const int sl = (threadIdx.x & 31) % sw;      // lane inside that sub-warp
  const uint32_t smask = sw >= 32 ? 0xffffffffu : (((1u << sw) - 1u) << (((threadIdx.x & 31) / sw) * sw));
  const int nchunk = d >> 2;
  const int g16 = (d + 511) / 512;
  for (int i = threadIdx.x; i < groups * PB_N_DISTS * d; i += kBwdThreads) dts[i] = 0.f;
  __syncthreads();
  float* my_dt = dts + (size_t)g * PB_N_DISTS * d;

  const int64_t per_cta = (n_nodes + gridDim.x - 1) / gridDim.x;
  const int64_t first = (int64_t)blockIdx.x * per_cta;
  const int64_t last = first + per_cta < n_nodes ? first + per_cta : n_nodes;

  constexpr int kBatch = 8;   // divides the shuffle width (16 or 32)
  struct NodeData {
    float4 xu[CPT], acc[CPT];
    int beg, end;
    int4 rec;
  };
  auto load_rows = [&](int64_t u, NodeData& nd) {
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = tc + tpn * j;
      if (c < nchunk) {
        nd.xu[j] = ldg4(x + (size_t)u * d + 4 * c);
        nd.acc[j] = load_grad4<BF16>(d_a, (size_t)u * ldda + (size_t)n_rel * d + 4 * c);  // root branch
        if (gy_res) {
          const float4 r = ldg4(gy_res + (size_t)u * d + 4 * c);  // residual branch
          nd.acc[j].x += r.x; nd.acc[j].y += r.y; nd.acc[j].z += r.z; nd.acc[j].w += r.w;
        }
      }
    }
    nd.beg = __ldg(out_ptr + u);
    nd.end = __ldg(out_ptr + u + 1);
  };
  auto load_recs = [&](int base, int end) {
    return base + sl < end ? __ldg(out_rec + base + sl) : make_int4(0, 0, 0, 0);
  };

  NodeData cur;
  int64_t u = first + g;
  if (u < last) {
    load_rows(u, cur);
    cur.rec = load_recs(cur.beg, cur.end);
  }
  while (u < last) {
    const int64_t un = u + groups;
    NodeData nxt;
    if (un < last) load_rows(un, nxt);
    int base = cur.beg;
    int4 my_rec = cur.rec;
    for (int b0 = cur.beg; b0 < cur.end; b0 += kBatch) {
      if (b0 - base == sw) {  // uniform across the node's threads (degree > shuffle width only)
        base = b0;
        my_rec = load_recs(base, cur.end);
      }
      const int nb = cur.end - b0 < kBatch ? cur.end - b0 : kBatch;
      float4 dh[kBatch][CPT];
      uint32_t nibs[kBatch];
#pragma unroll
      for (int i = 0; i < kBatch; ++i) {   // issue every gather of the batch before any of them is consumed
        const int dst = __shfl_sync(smask, my_rec.x, b0 - base + i, sw);
        const int meta = __shfl_sync(smask, my_rec.y, b0 - base + i, sw);
        const uint32_t eid = (uint32_t)__shfl_sync(smask, my_rec.z, b0 - base + i, sw);
        nibs[i] = 0xFFFFFFFFu;
        if (i < nb) {
          const size_t off = (size_t)dst * ldda + (size_t)(meta & 0xff) * d;
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            const int c = tc + tpn * j;
            if (c < nchunk) dh[i][j] = load_grad4<BF16>(d_a, off + 4 * c);
          }
          if constexpr (DROPOUT) {
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
              const int c = tc + tpn * j;
              if (c < nchunk) w |= keep_nibble(keep_bits, g16, eid, c) << (4 * j);
            }
            nibs[i] = w;
          }
        }
      }
      if (b0 == cur.beg && un < last) nxt.rec = load_recs(nxt.beg, nxt.end);   // next node's records, behind the gathers
#pragma unroll
      for (int i = 0; i < kBatch; ++i) {
        const int meta = __shfl_sync(smask, my_rec.y, b0 - base + i, sw);
        const int cnt = __shfl_sync(smask, my_rec.w, b0 - base + i, sw);
        if (i < nb) {
          const int dist = meta >> 8;
          // d(mean)/d(sum) = 1/|segment| (one reciprocal per edge), times the dropout scale
          float coef = cnt > 1 ? 1.0f / (float)cnt : 1.0f;
          if constexpr (DROPOUT) coef *= keep_scale;
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            const int c = tc + tpn * j;
            if (c < nchunk) {
              const float4 t = ldg4(table + (size_t)dist * d + 4 * c);
              const float4 xv = cur.xu[j];
              const uint32_t nib = nibs[i] >> (4 * j);
              float4 ds = dh[i][j];
              ds.x = ((nib & 1u) && xv.x * t.x > 0.f) ? ds.x * coef : 0.f;
              ds.y = ((nib & 2u) && xv.y * t.y > 0.f) ? ds.y * coef : 0.f;
              ds.z = ((nib & 4u) && xv.z * t.z > 0.f) ? ds.z * coef : 0.f;
              ds.w = ((nib & 8u) && xv.w * t.w > 0.f) ? ds.w * coef : 0.f;
              cur.acc[j].x += ds.x * t.x; cur.acc[j].y += ds.y * t.y;
              cur.acc[j].z += ds.z * t.z; cur.acc[j].w += ds.w * t.w;
              float4* slot = reinterpret_cast<float4*>(my_dt + (size_t)dist * d + 4 * c);
              float4 acc_t = *slot;
              acc_t.x += ds.x * xv.x; acc_t.y += ds.y * xv.y; acc_t.z += ds.z * xv.z; acc_t.w += ds.w * xv.w;
              *slot = acc_t;
            }
          }
        }
      }
    }
    if (cur.beg == cur.end && un < last) nxt.rec = load_recs(nxt.beg, nxt.end);   // isolated node: no batch ran
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = tc + tpn * j;
      if (c < nchunk) st_stream4(gx + (size_t)u * d + 4 * c, cur.acc[j]);
    }
    cur = nxt;
    u = un;
  }
  __syncthreads();
  float* out = dt_partials + (size_t)blockIdx.x * PB_N_DISTS * d;
  for (int i = threadIdx.x; i < PB_N_DISTS * d; i += kBwdThreads) {
    float s = 0.f;
    for (int gg = 0; gg < groups; ++gg) s += dts[(size_t)gg * PB_N_DISTS * d + i];
    out[i] = s;
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

extern "C" int pb_edge_table_bwd(const float* dtable_partials, int32_t n_partials, int32_t d, float* g_nn_weight,
                                 float* g_nn_bias, pb_stream_t stream) {
  PB_REQUIRE(dtable_partials && g_nn_weight && g_nn_bias && n_partials > 0 && d > 0, "pb_edge_table_bwd: bad arguments");
  edge_table_bwd_kernel<<<(d + 31) / 32, dim3(32, 32), 0, as_stream(stream)>>>(dtable_partials, n_partials, d,
                                                                               g_nn_weight, g_nn_bias);
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
  dropout_bits_kernel<<<grid, 256, 0, as_stream(stream)>>>(n_edges, d, dropout_thresh(p_drop), seed,
                                                           reinterpret_cast<uint16_t*>(keep_bits));
  PB_LAUNCH_CHECK();
  return PB_OK;
}

template <bool BF16, bool DROP>
static int launch_agg_fwd(const pb_csr_t* g, const float* x, int d, const float* table, void* a_hi, void* a_lo,
                          int64_t lda, const uint16_t* bits, float scale, cudaStream_t st) {
  const int cpl = (d + 127) / 128;
  const bool exact = d % 128 == 0 && (cpl == 1 || cpl == 2 || cpl == 4 || cpl == 8);
  const int threads = 256;
  const int64_t want = (g->n_nodes * 32 + threads - 1) / threads;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count() * 8 * 4));
#define PB_AGG_FWD(CPL, EX)                                                                                      \
  agg_fwd_kernel<BF16, DROP, CPL, EX><<<grid, threads, 0, st>>>(g->in_ptr, g->in_edge, g->in_eid, x, table, a_hi, \
                                                                a_lo, lda, g->n_nodes, d, g->n_relations, bits, scale)
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

extern "C" int pb_agg_fwd(const pb_csr_t* csr, const float* x, int32_t d, const float* table, void* a_hi, void* a_lo,
                          int64_t lda, int32_t dtype, const void* keep_bits, float p_drop, pb_stream_t stream) {
  int rc = check_csr(csr, d, "pb_agg_fwd");
  if (rc) return rc;
  PB_REQUIRE(x && table && a_hi, "pb_agg_fwd: null pointer");
  PB_REQUIRE(dtype == PB_BF16 || (dtype == PB_F32 && a_lo), "pb_agg_fwd: PB_F32 needs a_lo");
  PB_REQUIRE(lda >= (int64_t)(csr->n_relations + 1) * d && lda % 8 == 0, "pb_agg_fwd: bad lda");
  PB_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "pb_agg_fwd: p_drop out of range");
  PB_REQUIRE(!keep_bits || csr->in_eid, "pb_agg_fwd: dropout needs in_eid");
  PB_REQUIRE(p_drop == 0.f || keep_bits, "pb_agg_fwd: p_drop > 0 needs keep_bits (pb_dropout_bits)");
  cudaStream_t st = as_stream(stream);
  const uint16_t* bits = p_drop > 0.f ? reinterpret_cast<const uint16_t*>(keep_bits) : nullptr;
  const float scale = 1.f / (1.f - p_drop);
  if (dtype == PB_BF16)
    return bits ? launch_agg_fwd<true, true>(csr, x, d, table, a_hi, a_lo, lda, bits, scale, st)
                : launch_agg_fwd<true, false>(csr, x, d, table, a_hi, a_lo, lda, bits, scale, st);
  return bits ? launch_agg_fwd<false, true>(csr, x, d, table, a_hi, a_lo, lda, bits, scale, st)
              : launch_agg_fwd<false, false>(csr, x, d, table, a_hi, a_lo, lda, bits, scale, st);
}

extern "C" int32_t pb_agg_bwd_num_partials(void) { return kBwdPartials; }

template <bool BF16, bool DROP>
static int launch_agg_bwd(const pb_csr_t* g, const float* x, int d, const float* table, const void* d_a, int64_t ldda,
                          const float* gy_res, float* gx, float* dtp, const uint16_t* bits, float scale,
                          cudaStream_t st) {
  int tpn = 1;                                   // threads per node: largest power of two <= min(128, d/4)
  while (tpn * 2 <= std::min(kBwdThreads, d / 4)) tpn *= 2;
  const int groups = kBwdThreads / tpn;
  const int cpt = (d / 4 + tpn - 1) / tpn;
  const size_t smem = (size_t)groups * PB_N_DISTS * d * sizeof(float);
#define PB_AGG_BWD(CPT)                                                                                          \
  do {                                                                                                           \
    PB_CUDA(cudaFuncSetAttribute(agg_bwd_kernel<BF16, DROP, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                 (int)smem));                                                                    \
    agg_bwd_kernel<BF16, DROP, CPT><<<kBwdPartials, kBwdThreads, smem, st>>>(                                    \
        g->out_ptr, reinterpret_cast<const int4*>(g->out_rec), x, table, d_a, ldda, gy_res, gx, dtp, g->n_nodes, \
        d, g->n_relations, tpn, bits, scale);                                                                    \
  } while (0)
  if (cpt <= 1) PB_AGG_BWD(1);
  else PB_AGG_BWD(2);
#undef PB_AGG_BWD
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_agg_bwd(const pb_csr_t* csr, const float* x, int32_t d, const float* table, const void* d_a,
                          int64_t ldda, int32_t dtype, const float* gy_res, float* gx, float* dtable_partials,
                          const void* keep_bits, float p_drop, pb_stream_t stream) {
  int rc = check_csr(csr, d, "pb_agg_bwd");
  if (rc) return rc;
  PB_REQUIRE(x && table && d_a && gx && dtable_partials, "pb_agg_bwd: null pointer");
  PB_REQUIRE(dtype == PB_BF16 || dtype == PB_F32, "pb_agg_bwd: bad dtype");
  PB_REQUIRE(ldda >= (int64_t)(csr->n_relations + 1) * d && ldda % 8 == 0, "pb_agg_bwd: bad ldda");
  PB_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "pb_agg_bwd: p_drop out of range");
  PB_REQUIRE(p_drop == 0.f || keep_bits, "pb_agg_bwd: p_drop > 0 needs keep_bits (pb_dropout_bits)");
  cudaStream_t st = as_stream(stream);
  const uint16_t* bits = p_drop > 0.f ? reinterpret_cast<const uint16_t*>(keep_bits) : nullptr;
  const float scale = 1.f / (1.f - p_drop);
  if (dtype == PB_BF16)
    return bits ? launch_agg_bwd<true, true>(csr, x, d, table, d_a, ldda, gy_res, gx, dtable_partials, bits, scale, st)
                : launch_agg_bwd<true, false>(csr, x, d, table, d_a, ldda, gy_res, gx, dtable_partials, bits, scale, st);
  return bits ? launch_agg_bwd<false, true>(csr, x, d, table, d_a, ldda, gy_res, gx, dtable_partials, bits, scale, st)
              : launch_agg_bwd<false, false>(csr, x, d, table, d_a, ldda, gy_res, gx, dtable_partials, bits, scale, st);
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
