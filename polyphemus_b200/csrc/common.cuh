// Shared helpers for libpolyphemus_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <algorithm>

#include "../../include/polyphemus_b200.h"

namespace pb {

// ------------------------------------------------------------------ error plumbing (no exceptions over the ABI)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define PB_CUDA(expr)                                                        \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) return pb::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define PB_REQUIRE(cond, ...)              \
  do {                                     \
    if (!(cond)) {                         \
      pb::set_error(__VA_ARGS__);          \
      return PB_ERR_INVALID;               \
    }                                      \
  } while (0)

#define PB_LAUNCH_CHECK() PB_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(pb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // cached multiprocessor count of the current device

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// tensor-core aggregation backward (agg_bwd_tc.cu), dispatched by pb_agg_bwd_fused
bool agg_bwd_tc_eligible(int d, int dtype);
int agg_bwd_tc_ctas(int64_t n_nodes);
int agg_bwd_tc_launch(const pb_csr_t* g, const void* x, int d, const float* table, const void* d_a, int64_t ldda,
                      const void* gy_res, void* gx, float* partials, const uint16_t* bits, float scale, bool act_bf16,
                      cudaStream_t st);

// record kinds of the fused backward's stream (pb_csr_bwd_stream; layout next to bwd_stream_kernel in aggregate.cu)
enum : int { kRecEdge = 0, kRecX = 1, kRecRoot = 2, kRecRes = 3, kRecNop = 4, kRecLast = 8 };

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// streaming (read-once) 16-byte load / store: keep L1 for the gathered rows and the edge table
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_stream2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream2(void* p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));   // exact: bf16 = top half of fp32
}

// ------------------------------------------------------------------ activation storage (fp32 or bf16 rows)
// Node features x / y, the pre-BatchNorm output `out` and their gradients travel between the kernels of a stack either
// as fp32 (the API dtype, PB_F32) or as bf16 (PB_BF16: the throughput mode's stacks keep them in 16 bits, as the
// reference does under fp16 autocast). Arithmetic is fp32 either way; `elem` is an element offset, 4 elements a call.
template <bool ABF>
__device__ __forceinline__ float4 act_ld4(const void* base, size_t elem) {
  if constexpr (ABF) {
    const uint2 p = __ldg(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(base) + elem));
    const float2 a = unpack_bf16x2(p.x), b = unpack_bf16x2(p.y);
    return make_float4(a.x, a.y, b.x, b.y);
  } else {
    return ldg4(static_cast<const float*>(base) + elem);
  }
}
template <bool ABF>
__device__ __forceinline__ float4 act_ld4_stream(const void* base, size_t elem) {
  if constexpr (ABF) {
    const uint2 p = ld_stream2(static_cast<const __nv_bfloat16*>(base) + elem);
    const float2 a = unpack_bf16x2(p.x), b = unpack_bf16x2(p.y);
    return make_float4(a.x, a.y, b.x, b.y);
  } else {
    return ld_stream4(static_cast<const float*>(base) + elem);
  }
}
template <bool ABF>
__device__ __forceinline__ void act_st4_stream(void* base, size_t elem, float4 v) {
  if constexpr (ABF)
    st_stream2(static_cast<__nv_bfloat16*>(base) + elem, make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w)));
  else
    st_stream4(static_cast<float*>(base) + elem, v);
}
__device__ __forceinline__ float act_ld1(const void* base, size_t elem, bool abf) {
  return abf ? __bfloat162float(static_cast<const __nv_bfloat16*>(base)[elem]) : static_cast<const float*>(base)[elem];
}

// TF32 split used by the PB_F32 tensor-core mode: hi keeps the top 19 bits (sign, 8 exp, 10 mantissa),
// lo = v - hi is exact in fp32 and is itself truncated to TF32 by the tensor core (error <= 2^-21 |v|).
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_lo(float v, float hi) { return v - hi; }

// ------------------------------------------------------------------ counter-based dropout RNG (stateless)
// keep-mask of GCL.message's dropout (model.py:133). One 64-bit hash per (edge, 4-channel chunk): the SplitMix64
// output function (Stafford "Mix13" finaliser, the generator behind java.util.SplittableRandom) applied to the
// counter seed + GOLDEN * (1 + (eid << 32 | chunk)); the four 16-bit lanes decide the four channels.
// ~16 integer instructions per chunk — a Philox4x32-10 call costs ~5x that and made agg_fwd issue-bound.
__host__ __device__ inline uint64_t splitmix64_mix(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t dropout_bits(uint64_t seed, uint32_t eid, uint32_t chunk) {
  const uint64_t ctr = ((uint64_t)eid << 32) | (uint64_t)chunk;
  return splitmix64_mix(seed + 0x9E3779B97F4A7C15ull * (ctr + 1ull));
}
// keep channel i of the chunk iff its 16-bit lane >= thresh16 (= round(p * 65536))
__host__ __device__ inline void dropout_keep4(uint64_t seed, uint32_t eid, uint32_t chunk, uint32_t thresh16,
                                              bool keep[4]) {
  const uint64_t r = dropout_bits(seed, eid, chunk);
#pragma unroll
  for (int i = 0; i < 4; ++i) keep[i] = (uint32_t)((r >> (16 * i)) & 0xFFFFu) >= thresh16;
}
__host__ __device__ inline uint32_t dropout_thresh(float p) {
  double t = (double)p * 65536.0 + 0.5;
  return t >= 65535.0 ? 65535u : (uint32_t)t;
}
// scale that keeps the expectation exact for the quantised probability thresh16 / 65536
__host__ __device__ inline float dropout_scale(uint32_t thresh16) { return 65536.0f / (float)(65536u - thresh16); }

}  // namespace pb
