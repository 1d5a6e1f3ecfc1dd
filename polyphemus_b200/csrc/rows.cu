// Row permutation into / out of the structured node layout, fused with the activation-storage conversion.
//
// A GCN stack runs on rows sorted by track relation and padded per group to the GEMM tile; entering it is
// xp[pos[v]] = convert(x[v]) with the padding rows zeroed, leaving it y[v] = convert(xp[pos[v]]), and their gradients are
// each other. In ATen that is convert + zeros + index_copy_ / index_select + convert (5 kernels, the matrix moved three
// times); here one pass each: a warp per node row, 16-byte accesses. pos is injective, so there are no conflicts.
#include "common.cuh"

namespace pb {

template <bool SRC_BF, bool DST_BF>
__device__ __forceinline__ void copy_row(const void* src, size_t s_off, void* dst, size_t d_off, int d, int lane) {
  for (int c = 4 * lane; c < d; c += 128) {
    const float4 v = act_ld4_stream<SRC_BF>(src, s_off + c);
    act_st4_stream<DST_BF>(dst, d_off + c, v);
  }
}

// SCATTER: dst[pos[v]] = src[v] (v < n), else GATHER: dst[v] = src[pos[v]]
template <bool SRC_BF, bool DST_BF, bool SCATTER>
__global__ void __launch_bounds__(256) rows_permute_kernel(const void* __restrict__ src, const long long* __restrict__ pos,
                                                           int64_t n, int d, void* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < n; v += n_warps) {
    const long long p = __ldg(pos + v);
    if (SCATTER) copy_row<SRC_BF, DST_BF>(src, (size_t)v * d, dst, (size_t)p * d, d, lane);
    else copy_row<SRC_BF, DST_BF>(src, (size_t)p * d, dst, (size_t)v * d, d, lane);
  }
}

struct PadRanges { int n; long long begin[4], end[4]; };
template <bool DST_BF>
__global__ void rows_zero_pad_kernel(void* __restrict__ dst, int d, const PadRanges pr) {
  const int nchunk = d >> 2;
  for (int g = 0; g < pr.n; ++g) {
    const long long rows = pr.end[g] - pr.begin[g];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * nchunk; i += (long long)gridDim.x * blockDim.x)
      act_st4_stream<DST_BF>(dst, (size_t)(pr.begin[g] + i / nchunk) * d + 4 * (size_t)(i % nchunk), make_float4(0.f, 0.f, 0.f, 0.f));
  }
}

template <bool SCATTER>
static int launch_permute(const void* src, int sdt, const int64_t* pos, int64_t n, int d, void* dst, int ddt, cudaStream_t st) {
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n * 32 + 255) / 256, (int64_t)sm_count() * 16));
  const long long* p = reinterpret_cast<const long long*>(pos);
#define PB_PERM(S, D) rows_permute_kernel<S, D, SCATTER><<<grid, 256, 0, st>>>(src, p, n, d, dst)
  if (sdt == PB_BF16) { if (ddt == PB_BF16) PB_PERM(true, true); else PB_PERM(true, false); }
  else { if (ddt == PB_BF16) PB_PERM(false, true); else PB_PERM(false, false); }
#undef PB_PERM
  PB_LAUNCH_CHECK();
  return PB_OK;
}

}  // namespace pb

using namespace pb;

static int check_rows(const void* src, const int64_t* pos, void* dst, int64_t n, int d, int sdt, int ddt, const char* who) {
  PB_REQUIRE(src && pos && dst && n >= 0, "%s: bad arguments", who);
  PB_REQUIRE(d > 0 && d % 8 == 0, "%s: d=%d must be a multiple of 8", who, d);
  PB_REQUIRE((sdt == PB_F32 || sdt == PB_BF16) && (ddt == PB_F32 || ddt == PB_BF16), "%s: bad dtype", who);
  return PB_OK;
}

extern "C" int pb_rows_scatter(const void* src, int32_t src_dtype, const int64_t* pos, int64_t n, int32_t d, void* dst,
                               int32_t dst_dtype, int64_t n_rows, const pb_groups_t* groups, pb_stream_t stream) {
  if (int rc = check_rows(src, pos, dst, n, d, src_dtype, dst_dtype, "pb_rows_scatter")) return rc;
  PB_REQUIRE(groups && groups->n_groups > 0 && groups->n_groups <= 4 && n_rows >= n, "pb_rows_scatter: needs the row groups");
  cudaStream_t st = as_stream(stream);
  PadRanges pr;
  pr.n = 0;
  long long n_pad = 0;
  for (int g = 0; g < groups->n_groups; ++g) {
    const long long b = groups->start[g] + groups->count[g];
    const long long e = g + 1 < groups->n_groups ? groups->start[g + 1] : (long long)n_rows;
    if (e > b) { pr.begin[pr.n] = b; pr.end[pr.n] = e; ++pr.n; n_pad += e - b; }
  }
  if (n_pad > 0) {
    const unsigned grid = (unsigned)std::min<long long>((n_pad * (d / 4) + 255) / 256, 64);
    if (dst_dtype == PB_BF16) rows_zero_pad_kernel<true><<<grid, 256, 0, st>>>(dst, d, pr);
    else rows_zero_pad_kernel<false><<<grid, 256, 0, st>>>(dst, d, pr);
    PB_LAUNCH_CHECK();
  }
  if (n == 0) return PB_OK;
  return launch_permute<true>(src, src_dtype, pos, n, d, dst, dst_dtype, st);
}

extern "C" int pb_rows_gather(const void* src, int32_t src_dtype, const int64_t* pos, int64_t n, int32_t d, void* dst,
                              int32_t dst_dtype, pb_stream_t stream) {
  if (int rc = check_rows(src, pos, dst, n, d, src_dtype, dst_dtype, "pb_rows_gather")) return rc;
  if (n == 0) return PB_OK;
  return launch_permute<false>(src, src_dtype, pos, n, d, dst, dst_dtype, as_stream(stream));
}
