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

// exp of a softmax term (argument <= 0). bf16 logits: ex2.approx on x * log2(e), relative error ~2^-21 — far below the
// rounding of the logits themselves and of the bf16 gradient, 2 instructions instead of ~12. fp32 logits: expf.
template <bool BF16>
__device__ __forceinline__ float ce_exp(float x) {
  if constexpr (BF16) return __expf(x);
  else return expf(x);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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
        s += ce_exp<BF16>(v[i][c][j] - m);
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
      if (live) {                                                // warp-uniform
#pragma unroll
        for (int j = 0; j < V; ++j) o[j] = (ce_exp<BF16>(v[i][c][j] - l[i]) - (col + j == tgt[i] ? 1.f : 0.f)) * rg[i];
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) o[j] = 0.f;
      }
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

// ---- several heads in the column blocks of one matrix: one warp per row, the whole row in one pass --------------
struct CeSegments {
  int n;
  int col0[PB_CE_MAX_SEGMENTS + 1];            // block s = columns [col0[s], col0[s+1])
  int ignore[PB_CE_MAX_SEGMENTS];
  const int* target[PB_CE_MAX_SEGMENTS];
};

// One warp per RPW = 4 consecutive rows. The (row, segment) targets / log-sum-exps / row gradients are fetched
// lane-parallel (one coalesced load each, then shuffles), every live (row, segment) block is loaded as raw 16-byte
// chunks before the first use (CPS chunks per lane and segment: width <= 32 * V * CPS), and ignored blocks are never
// read. Segment loops are unrolled over the template segment count S, so the by-value struct stays in the constant
// bank. The backward writes whole rows (zeros for ignored blocks).
template <bool BF16> struct CeRaw { using type = float4; };
template <> struct CeRaw<true> { using type = uint4; };

template <bool BF16>
__device__ __forceinline__ void ce_unpack(const typename CeRaw<BF16>::type& raw, float (&o)[BF16 ? 8 : 4]) {
  if constexpr (BF16) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[2 * j] = __uint_as_float(w[j] << 16);
      o[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    }
  } else {
    o[0] = raw.x, o[1] = raw.y, o[2] = raw.z, o[3] = raw.w;
  }
}

template <bool BF16, int CPS, int RPW, int S, bool BACKWARD>
__global__ void __launch_bounds__(kCeThreads)
ce_rows_kernel(const void* __restrict__ logits, long ld, long rows, CeSegments seg, float* __restrict__ nll,
               float* __restrict__ lse, const float* __restrict__ row_grad, void* __restrict__ grad, long ldg) {
  constexpr int V = BF16 ? 8 : 4;
  constexpr uint32_t kFull = 0xffffffffu;
  static_assert(RPW * S <= 32, "one lane per (row, segment)");
  using Raw = typename CeRaw<BF16>::type;
  const int lane = threadIdx.x & 31;
  const long r0 = ((static_cast<long>(blockIdx.x) * kCeThreads + threadIdx.x) >> 5) * RPW;
  // lane = s * RPW + i holds the scalars of (row i, segment s)
  const int my_s = lane / RPW, my_i = lane % RPW;
  int my_tgt = 0;
  bool my_live = false;
  float my_l = 0.f, my_g = 0.f;
#pragma unroll
  for (int s = 0; s < S; ++s)
    if (my_s == s && r0 + my_i < rows) {
      my_tgt = __ldg(seg.target[s] + r0 + my_i);
      my_live = my_tgt != seg.ignore[s];
      if constexpr (BACKWARD)
        if (my_live) {
          my_l = __ldg(lse + static_cast<long>(s) * rows + r0 + my_i);
          my_g = __ldg(row_grad + static_cast<long>(s) * rows + r0 + my_i);
        }
    }
  const uint32_t live_mask = __ballot_sync(kFull, my_live);
  Raw raw[RPW][S][CPS];
#pragma unroll
  for (int i = 0; i < RPW; ++i)
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
      for (int c = 0; c < CPS; ++c) {
        const int col = seg.col0[s] + (c * 32 + lane) * V;
        if (((live_mask >> (s * RPW + i)) & 1u) && col < seg.col0[s + 1]) {
          if constexpr (BF16)
            raw[i][s][c] = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(logits) + (r0 + i) * ld + col));
          else
            raw[i][s][c] = ldg4(static_cast<const float*>(logits) + (r0 + i) * ld + col);
        } else {
          if constexpr (BF16) raw[i][s][c] = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);   // -inf
          else raw[i][s][c] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
      }
  constexpr float kLog2e = 1.4426950408889634f;
  if constexpr (!BACKWARD) {
    // the target logit of this lane's (row, segment): one scalar load (the row is being fetched anyway)
    float my_xt = 0.f;
#pragma unroll
    for (int s = 0; s < S; ++s)
      if (my_s == s && my_live) {
        const long off = (r0 + my_i) * ld + seg.col0[s] + my_tgt;
        if constexpr (BF16) my_xt = __bfloat162float(static_cast<const __nv_bfloat16*>(logits)[off]);
        else my_xt = static_cast<const float*>(logits)[off];
      }
    float my_m = 0.f, my_sum = 1.f;
#pragma unroll
    for (int i = 0; i < RPW; ++i)
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int owner = s * RPW + i;
        if (!((live_mask >> owner) & 1u)) continue;                // warp-uniform
        float v[CPS][V];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < CPS; ++c) {
          ce_unpack<BF16>(raw[i][s][c], v[c]);
#pragma unroll
          for (int j = 0; j < V; ++j) m = fmaxf(m, v[c][j]);
        }
        m = warp_max(m);
        float sum = 0.f;
        if constexpr (BF16) {                                      // ex2.approx(v log2e - m log2e): 2 instructions a term
          const float ms = m * kLog2e;
#pragma unroll
          for (int c = 0; c < CPS; ++c)
#pragma unroll
            for (int j = 0; j < V; ++j) sum += ex2_approx(fmaf(v[c][j], kLog2e, -ms));
        } else {
#pragma unroll
          for (int c = 0; c < CPS; ++c)
#pragma unroll
            for (int j = 0; j < V; ++j) sum += expf(v[c][j] - m);
        }
        sum = warp_sum(sum);
        if (lane == owner) my_m = m, my_sum = sum;
      }
#pragma unroll
    for (int s = 0; s < S; ++s)
      if (my_s == s && r0 + my_i < rows) {
        const float l = my_live ? my_m + logf(my_sum) : 0.f;       // one logf per lane, not per block
        lse[static_cast<long>(s) * rows + r0 + my_i] = l;
        nll[static_cast<long>(s) * rows + r0 + my_i] = my_live ? l - my_xt : 0.f;
      }
  } else {
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      const long r = r0 + i;
      if (r >= rows) break;                                        // warp-uniform
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int owner = s * RPW + i;
        const bool live = (live_mask >> owner) & 1u;               // warp-uniform
        const int tgt = __shfl_sync(kFull, my_tgt, owner);
        const float l = __shfl_sync(kFull, my_l, owner), g = __shfl_sync(kFull, my_g, owner);
#pragma unroll
        for (int c = 0; c < CPS; ++c) {
          const int rel = (c * 32 + lane) * V;
          const int col = seg.col0[s] + rel;
          if (col >= seg.col0[s + 1]) continue;
          float o[V];
          if (live) {                                              // ignored blocks cost a store only
            float v[V];
            ce_unpack<BF16>(raw[i][s][c], v);
            if constexpr (BF16) {
              const float ls = l * kLog2e;
#pragma unroll
              for (int j = 0; j < V; ++j) o[j] = ex2_approx(fmaf(v[j], kLog2e, -ls)) * g;
            } else {
#pragma unroll
              for (int j = 0; j < V; ++j) o[j] = expf(v[j] - l) * g;
            }
            const int hit = tgt - rel;                              // the one-hot term lands in at most one lane
#pragma unroll
            for (int j = 0; j < V; ++j) o[j] -= (hit == j) ? g : 0.f;   // select, not a branch
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j) o[j] = 0.f;
          }
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
  }
}

int ce_rows_check(const char* who, const void* logits, int64_t ld, int32_t dtype, int64_t rows, int32_t n_segments,
                  const int32_t* widths, const int32_t* const* targets, CeSegments* seg, int* chunks) {
  PB_REQUIRE(dtype == PB_BF16 || dtype == PB_F32, "%s: dtype %d", who, dtype);
  PB_REQUIRE(n_segments >= 1 && n_segments <= PB_CE_MAX_SEGMENTS && widths && targets, "%s: 1..%d segments", who,
             PB_CE_MAX_SEGMENTS);
  const int v = dtype == PB_BF16 ? 8 : 4;
  seg->n = n_segments;
  seg->col0[0] = 0;
  for (int s = 0; s < n_segments; ++s) {
    PB_REQUIRE(widths[s] > 0 && widths[s] % v == 0 && targets[s], "%s: segment %d width %d must be a multiple of %d", who, s,
               widths[s], v);
    seg->col0[s + 1] = seg->col0[s] + widths[s];
    seg->target[s] = targets[s];
  }
  const int cols = seg->col0[n_segments];
  PB_REQUIRE(rows >= 0 && ld >= cols && ld % v == 0 && reinterpret_cast<uintptr_t>(logits) % 16 == 0,
             "%s: row stride %lld / alignment", who, static_cast<long long>(ld));
  int widest = 0;
  for (int s = 0; s < n_segments; ++s) widest = std::max(widest, widths[s]);
  const int need = (widest + 32 * v - 1) / (32 * v);
  PB_REQUIRE(need <= 2, "%s: at most %d columns per segment", who, 2 * 32 * v);
  *chunks = need;
  return PB_OK;
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

#define PB_CE_ROWS_LAUNCH(BF, CPS, S, BWD, ...) \
  ce_rows_kernel<BF, CPS, 4, S, BWD><<<grid, kCeThreads, 0, as_stream(stream)>>>(__VA_ARGS__)
#define PB_CE_ROWS_BY_SEG(BF, CPS, BWD, ...)                             \
  do {                                                                   \
    if (n_segments == 1) PB_CE_ROWS_LAUNCH(BF, CPS, 1, BWD, __VA_ARGS__); \
    else if (n_segments == 2) PB_CE_ROWS_LAUNCH(BF, CPS, 2, BWD, __VA_ARGS__); \
    else if (n_segments == 3) PB_CE_ROWS_LAUNCH(BF, CPS, 3, BWD, __VA_ARGS__); \
    else PB_CE_ROWS_LAUNCH(BF, CPS, 4, BWD, __VA_ARGS__);                \
  } while (0)
#define PB_CE_ROWS_DISPATCH(BWD, ...)                                            \
  do {                                                                           \
    const unsigned grid = static_cast<unsigned>(((rows + 3) / 4 + 7) / 8);       \
    if (dtype == PB_BF16) {                                                      \
      if (chunks == 1) PB_CE_ROWS_BY_SEG(true, 1, BWD, __VA_ARGS__);             \
      else PB_CE_ROWS_BY_SEG(true, 2, BWD, __VA_ARGS__);                         \
    } else {                                                                     \
      if (chunks == 1) PB_CE_ROWS_BY_SEG(false, 1, BWD, __VA_ARGS__);            \
      else PB_CE_ROWS_BY_SEG(false, 2, BWD, __VA_ARGS__);                        \
    }                                                                            \
  } while (0)

extern "C" int pb_ce_rows_fwd(const void* logits, int64_t ld, int32_t dtype, int64_t rows, int32_t n_segments,
                              const int32_t* widths, const int32_t* const* targets, const int32_t* ignore_index,
                              float* nll, float* lse, pb_stream_t stream) {
  CeSegments seg{};
  int chunks = 0;
  if (int rc = ce_rows_check("pb_ce_rows_fwd", logits, ld, dtype, rows, n_segments, widths, targets, &seg, &chunks)) return rc;
  PB_REQUIRE(ignore_index && nll && lse, "pb_ce_rows_fwd: null pointer");
  for (int s = 0; s < n_segments; ++s) seg.ignore[s] = ignore_index[s];
  if (rows == 0) return PB_OK;
  PB_CE_ROWS_DISPATCH(false, logits, ld, rows, seg, nll, lse, nullptr, nullptr, 0);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_ce_rows_bwd(const void* logits, int64_t ld, int32_t dtype, int64_t rows, int32_t n_segments,
                              const int32_t* widths, const int32_t* const* targets, const int32_t* ignore_index,
                              const float* lse, const float* row_grad, void* grad, int64_t ldg, pb_stream_t stream) {
  CeSegments seg{};
  int chunks = 0;
  if (int rc = ce_rows_check("pb_ce_rows_bwd", logits, ld, dtype, rows, n_segments, widths, targets, &seg, &chunks)) return rc;
  PB_REQUIRE(ignore_index && lse && row_grad && grad, "pb_ce_rows_bwd: null pointer");
  PB_REQUIRE(ldg % (dtype == PB_BF16 ? 8 : 4) == 0 && ldg >= seg.col0[n_segments] && reinterpret_cast<uintptr_t>(grad) % 16 == 0,
             "pb_ce_rows_bwd: grad stride %lld / alignment", static_cast<long long>(ldg));
  for (int s = 0; s < n_segments; ++s) seg.ignore[s] = ignore_index[s];
  if (rows == 0) return PB_OK;
  PB_CE_ROWS_DISPATCH(true, logits, ld, rows, seg, nullptr, const_cast<float*>(lse), row_grad, grad, ldg);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
