// Chord embedding of the content encoder (model.py:355-388): per node, 15 token slots, each a (pitch id, duration id)
// pair; the reference embeds the one-hot tokens (Linear 131->d/2, Linear 99->d/2), batch-normalises, concatenates the
// 15 x d values and applies chord_encoder = Linear(15 d, d), ReLU.
//
// Every stage before the ReLU is linear in the one-hot tokens, so the chord pre-activation is a sum of 30 rows of a
// folded table  T[set, slot, token, :] = BN(emb)[token] @ W_chord[:, slot, half]^T  (built per step on the host side
// from the live parameters — 2 x 15 x 230 x d values, L2-resident):
//     chord[v] = relu(bias + sum_t T[set_v, t, pitch_v,t] + T[set_v, t, 131 + dur_v,t]),   set_v = is_drum[v].
// Forward is a gather-sum (one warp per node, 16-byte row gathers that hit L2/L1); the table gradient is
// onehot^T @ (g * relu') — a contraction over the nodes that runs on the deterministic split-K tcgen05 weight-gradient
// GEMM, fed by the operand-building kernel below. Replaces two [N*15, d/2] gathers, a [N, 15 d] concatenation and
// three 1-TFLOP GEMMs per step.
#include "common.cuh"

namespace pb {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float4 bf16x4_to_f4(uint2 p) {
  return make_float4(__uint_as_float(p.x << 16), __uint_as_float(p.x & 0xffff0000u), __uint_as_float(p.y << 16),
                     __uint_as_float(p.y & 0xffff0000u));
}

template <bool BF16, int CPL>
__global__ void __launch_bounds__(kThreads)
chord_embed_fwd_kernel(const int16_t* __restrict__ tokens, long tok_stride, int tok_offset, int n_slots,
                       const uint8_t* __restrict__ set_id, const void* __restrict__ tables, int vocab, int dur_off,
                       int d, const float* __restrict__ bias, float* __restrict__ out, long ldo, long n_nodes) {
  constexpr uint32_t kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int nchunk = d >> 2;
  const long v = (static_cast<long>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
  if (v >= n_nodes) return;
  const int n_ids = 2 * n_slots;                                   // <= 32 (checked by the host)
  const int my_id = lane < n_ids ? static_cast<int>(__ldg(tokens + v * tok_stride + tok_offset + lane)) : 0;
  const size_t set_base = static_cast<size_t>(__ldg(set_id + v) ? 1 : 0) * n_slots * vocab;
  float4 acc[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j)
    acc[j] = lane + 32 * j < nchunk ? ldg4(bias + 4 * (lane + 32 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
  for (int t = 0; t < n_slots; ++t) {
    int pid = __shfl_sync(kFull, my_id, 2 * t), did = __shfl_sync(kFull, my_id, 2 * t + 1);
    pid = min(max(pid, 0), dur_off - 1);
    did = min(max(did, 0), vocab - dur_off - 1);
    const size_t slot_base = set_base + static_cast<size_t>(t) * vocab;
    const size_t rp = (slot_base + pid) * d, rd = (slot_base + dur_off + did) * d;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      if (lane + 32 * j < nchunk) {
        const int c = 4 * (lane + 32 * j);
        float4 a, b;
        if constexpr (BF16) {
          const __nv_bfloat16* tb = static_cast<const __nv_bfloat16*>(tables);
          a = bf16x4_to_f4(__ldg(reinterpret_cast<const uint2*>(tb + rp + c)));
          b = bf16x4_to_f4(__ldg(reinterpret_cast<const uint2*>(tb + rd + c)));
        } else {
          const float* tb = static_cast<const float*>(tables);
          a = ldg4(tb + rp + c);
          b = ldg4(tb + rd + c);
        }
        acc[j].x += a.x + b.x; acc[j].y += a.y + b.y; acc[j].z += a.z + b.z; acc[j].w += a.w + b.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < CPL; ++j)
    if (lane + 32 * j < nchunk) {
      const float4 r = make_float4(fmaxf(acc[j].x, 0.f), fmaxf(acc[j].y, 0.f), fmaxf(acc[j].z, 0.f), fmaxf(acc[j].w, 0.f));
      st_stream4(out + v * ldo + 4 * (lane + 32 * j), r);
    }
}

// Operands of the table-gradient GEMM, one warp per node:
//   onehot [N, n_slots * vp]   1 at (slot t, pitch id) and (slot t, dur_off + dur id)
//   gcat   [N, 2 d]            g * 1[chord > 0] in the column block of the node's set, 0 in the other
// PB_BF16: both bf16. PB_F32: onehot fp32 (exact, no low part) and gcat as the TF32 hi/lo pair.
template <bool BF16>
__global__ void __launch_bounds__(kThreads)
chord_embed_bwd_prep_kernel(const int16_t* __restrict__ tokens, long tok_stride, int tok_offset, int n_slots,
                            const uint8_t* __restrict__ set_id, int dur_off, int vp, int d,
                            const float* __restrict__ chord, long ldc, const float* __restrict__ g, long ldg,
                            void* __restrict__ onehot, void* __restrict__ gcat_hi, float* __restrict__ gcat_lo,
                            long n_nodes) {
  constexpr uint32_t kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long v = (static_cast<long>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
  if (v >= n_nodes) return;
  const int n_ids = 2 * n_slots;
  const int my_id = lane < n_ids ? static_cast<int>(__ldg(tokens + v * tok_stride + tok_offset + lane)) : 0;
  const int set = __ldg(set_id + v) ? 1 : 0;
  const long oh_ld = static_cast<long>(n_slots) * vp;
  for (int t = 0; t < n_slots; ++t) {
    const int pid = __shfl_sync(kFull, my_id, 2 * t), did = dur_off + __shfl_sync(kFull, my_id, 2 * t + 1);
    if constexpr (BF16) {
      __nv_bfloat16* row = static_cast<__nv_bfloat16*>(onehot) + v * oh_ld + static_cast<long>(t) * vp;
      for (int c = 8 * lane; c < vp; c += 256) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c0 = c + 2 * q, c1 = c0 + 1;
          w[q] = ((c0 == pid || c0 == did) ? 0x3f80u : 0u) | ((c1 == pid || c1 == did) ? 0x3f800000u : 0u);
        }
        *reinterpret_cast<uint4*>(row + c) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else {
      float* row = static_cast<float*>(onehot) + v * oh_ld + static_cast<long>(t) * vp;
      for (int c = 4 * lane; c < vp; c += 128)
        *reinterpret_cast<float4*>(row + c) =
            make_float4((c == pid || c == did) ? 1.f : 0.f, (c + 1 == pid || c + 1 == did) ? 1.f : 0.f,
                        (c + 2 == pid || c + 2 == did) ? 1.f : 0.f, (c + 3 == pid || c + 3 == did) ? 1.f : 0.f);
    }
  }
  for (int c = 4 * lane; c < d; c += 128) {
    const float4 y = ldg4(chord + v * ldc + c);
    float4 gv = ld_stream4(g + v * ldg + c);
    gv.x = y.x > 0.f ? gv.x : 0.f; gv.y = y.y > 0.f ? gv.y : 0.f;
    gv.z = y.z > 0.f ? gv.z : 0.f; gv.w = y.w > 0.f ? gv.w : 0.f;
    const long own = v * 2 * d + static_cast<long>(set) * d + c, other = v * 2 * d + static_cast<long>(1 - set) * d + c;
    if constexpr (BF16) {
      __nv_bfloat16* o = static_cast<__nv_bfloat16*>(gcat_hi);
      const __nv_bfloat162 a = __floats2bfloat162_rn(gv.x, gv.y), b = __floats2bfloat162_rn(gv.z, gv.w);
      *reinterpret_cast<uint2*>(o + own) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
      *reinterpret_cast<uint2*>(o + other) = make_uint2(0u, 0u);
    } else {
      float* hi = static_cast<float*>(gcat_hi);
      const float4 h = make_float4(tf32_hi(gv.x), tf32_hi(gv.y), tf32_hi(gv.z), tf32_hi(gv.w));
      *reinterpret_cast<float4*>(hi + own) = h;
      *reinterpret_cast<float4*>(gcat_lo + own) = make_float4(gv.x - h.x, gv.y - h.y, gv.z - h.z, gv.w - h.w);
      *reinterpret_cast<float4*>(hi + other) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(gcat_lo + other) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

int chord_check(const char* who, int64_t n_nodes, int32_t n_slots, int64_t tok_stride, int32_t tok_offset, int32_t vocab,
                int32_t dur_off, int32_t d, int32_t dtype) {
  PB_REQUIRE(dtype == PB_BF16 || dtype == PB_F32, "%s: dtype %d", who, dtype);
  PB_REQUIRE(n_nodes >= 0 && n_slots > 0 && n_slots <= 16 && tok_offset >= 0 && tok_stride >= tok_offset + 2 * n_slots,
             "%s: %d slots at offset %d do not fit a token stride of %lld", who, n_slots, tok_offset,
             static_cast<long long>(tok_stride));
  PB_REQUIRE(vocab > dur_off && dur_off > 0, "%s: vocab %d / duration offset %d", who, vocab, dur_off);
  PB_REQUIRE(d >= 4 && d % 4 == 0 && d <= 1024, "%s: d=%d must be a multiple of 4, at most 1024", who, d);
  return PB_OK;
}

}  // namespace
}  // namespace pb

using namespace pb;

extern "C" int pb_chord_embed_fwd(const int16_t* tokens, int64_t tok_stride, int32_t tok_offset, int32_t n_slots,
                                  const uint8_t* set_id, const void* tables, int32_t dtype, int32_t vocab,
                                  int32_t dur_off, int32_t d, const float* bias, float* chord, int64_t ldc,
                                  int64_t n_nodes, pb_stream_t stream) {
  if (int rc = chord_check("pb_chord_embed_fwd", n_nodes, n_slots, tok_stride, tok_offset, vocab, dur_off, d, dtype)) return rc;
  PB_REQUIRE(tokens && set_id && tables && bias && chord && ldc >= d && ldc % 4 == 0, "pb_chord_embed_fwd: bad arguments");
  if (n_nodes == 0) return PB_OK;
  const unsigned grid = static_cast<unsigned>((n_nodes + 7) / 8);
  const int cpl = (d / 4 + 31) / 32;
#define PB_CHORD_FWD(BF, CPL)                                                                                       \
  chord_embed_fwd_kernel<BF, CPL><<<grid, kThreads, 0, as_stream(stream)>>>(tokens, tok_stride, tok_offset, n_slots, \
                                                                            set_id, tables, vocab, dur_off, d, bias, \
                                                                            chord, ldc, n_nodes)
  if (dtype == PB_BF16) {
    if (cpl <= 1) PB_CHORD_FWD(true, 1); else if (cpl <= 2) PB_CHORD_FWD(true, 2);
    else if (cpl <= 4) PB_CHORD_FWD(true, 4); else PB_CHORD_FWD(true, 8);
  } else {
    if (cpl <= 1) PB_CHORD_FWD(false, 1); else if (cpl <= 2) PB_CHORD_FWD(false, 2);
    else if (cpl <= 4) PB_CHORD_FWD(false, 4); else PB_CHORD_FWD(false, 8);
  }
#undef PB_CHORD_FWD
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_chord_embed_bwd_prep(const int16_t* tokens, int64_t tok_stride, int32_t tok_offset, int32_t n_slots,
                                       const uint8_t* set_id, int32_t dur_off, int32_t vocab_padded, int32_t d,
                                       const float* chord, int64_t ldc, const float* g, int64_t ldg, int32_t dtype,
                                       void* onehot, void* gcat_hi, void* gcat_lo, int64_t n_nodes,
                                       pb_stream_t stream) {
  if (int rc = chord_check("pb_chord_embed_bwd_prep", n_nodes, n_slots, tok_stride, tok_offset, vocab_padded, dur_off, d, dtype))
    return rc;
  PB_REQUIRE(tokens && set_id && chord && g && onehot && gcat_hi && (dtype == PB_BF16 || gcat_lo),
             "pb_chord_embed_bwd_prep: bad arguments");
  PB_REQUIRE(vocab_padded % 8 == 0 && ldc % 4 == 0 && ldg % 4 == 0, "pb_chord_embed_bwd_prep: strides / padded vocab");
  if (n_nodes == 0) return PB_OK;
  const unsigned grid = static_cast<unsigned>((n_nodes + 7) / 8);
  if (dtype == PB_BF16)
    chord_embed_bwd_prep_kernel<true><<<grid, kThreads, 0, as_stream(stream)>>>(
        tokens, tok_stride, tok_offset, n_slots, set_id, dur_off, vocab_padded, d, chord, ldc, g, ldg, onehot, gcat_hi,
        nullptr, n_nodes);
  else
    chord_embed_bwd_prep_kernel<false><<<grid, kThreads, 0, as_stream(stream)>>>(
        tokens, tok_stride, tok_offset, n_slots, set_id, dur_off, vocab_padded, d, chord, ldc, g, ldg, onehot, gcat_hi,
        static_cast<float*>(gcat_lo), n_nodes);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
