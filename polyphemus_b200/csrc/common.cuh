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
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// TF32 split used by the PB_F32 tensor-core mode: hi keeps the top 19 bits (sign, 8 exp, 10 mantissa),
// lo = v - hi is exact in fp32 and is itself truncated to TF32 by the tensor core (error <= 2^-21 |v|).
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_lo(float v, float hi) { return v - hi; }

// ------------------------------------------------------------------ Philox4x32-10 (counter based, stateless)
struct Philox {
  static constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
  }
  // counter = (c0,c1,c2,c3), key = (k0,k1)  ->  4 x u32
  __host__ __device__ static inline void run(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                             uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t h0, l0, h1, l1;
      mulhilo(kM0, c0, h0, l0);
      mulhilo(kM1, c2, h1, l1);
      uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += kW0; k1 += kW1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

// keep-mask bits for channels [4*chunk, 4*chunk+4) of edge `eid`: keep iff u32 >= thresh, thresh = p * 2^32
__host__ __device__ inline void dropout_keep4(uint64_t seed, uint32_t eid, uint32_t chunk, uint32_t thresh,
                                              bool keep[4]) {
  uint32_t r[4];
  Philox::run(eid, chunk, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
  for (int i = 0; i < 4; ++i) keep[i] = r[i] >= thresh;
}
__host__ __device__ inline uint32_t dropout_thresh(float p) {
  double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}

}  // namespace pb
