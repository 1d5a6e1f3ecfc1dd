// tcgen05 tensor-core contractions of the relational graph convolution (sm_100a only).
//
// Replaces the 7 N x d x d matmuls of GCL.forward (reference model.py:112,116) and the three matmul
// families autograd derives from them, each as ONE persistent, warp-specialised kernel:
//   fwd       out[M,d]  = A[M,K] . Wcat[K,d] + bias      (A K-major,  B = Wcat^T K-major)
//   bwd-data  dA[M,K]   = g[M,d] . Wcat^T                (A K-major,  B = Wcat   K-major)
//   bwd-wt    dWcat[K,d]= A^T . g  (split over nodes)    (A MN-major, B MN-major — no transposed copies)
//
// Structure (one CTA per SM; 192 threads, 320 in the bf16 kernel):
//   warp 0    TMA producer: cp.async.bulk.tensor (128B swizzle) into a multi-stage smem ring, mbarrier tx
//   warp 1    MMA issuer: one lane issues tcgen05.mma (A,B from smem descriptors, D in TMEM, fp32 accumulate),
//             tcgen05.commit releases smem stages and publishes the accumulator
//   warps 2-5 epilogue: tcgen05.ld the accumulator (one TMEM lane quadrant per warp), bias / convert, store
//   warps 6-9 (bf16 kernel, bf16 output): second epilogue group, the two groups split the tile's columns
// Two TMEM accumulator buffers let the epilogue of tile i overlap the main loop of tile i+1.
//
// Precision modes: PB_BF16 -> kind::f16 (bf16 operands), 1 MMA per k-step, tile 128x256x64;
//                  PB_F32  -> kind::tf32 on TF32 hi/lo splits, 3 MMAs per k-step (hi*hi + hi*lo + lo*hi),
//                             tile 128x128x32: fp32-grade results from the tensor cores.
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "mbar.cuh"
#include "tc.cuh"

namespace pb {

constexpr int kMaxSplits = 64;

struct GemmParams {
  CUtensorMap tm_a[2];   // hi, lo
  CUtensorMap tm_b[2];
  int64_t m, n;          // output extent
  int a_mn_major, b_mn_major;
  int num_m_tiles, num_n_tiles, num_splits;
  int k_blocks;
  int split_kb[kMaxSplits + 1];   // split s contracts over k-blocks [split_kb[s], split_kb[s+1])
  uint32_t idesc;
  // epilogue
  void* out;             // fp32 or bf16
  int64_t ldd;
  int64_t split_stride;  // elements between split-K partial outputs
  const float* bias;
  int out_bf16;
  // structured (grouped) weight selection: rows of group g use Wcat block g for the first remap_d contraction /
  // output columns and the shared blocks (offset +3*remap_d) for the rest
  int remap_mode;        // 0 none, 1 remap B's k coordinate (forward), 2 remap B's row coordinate (input gradient)
  int remap_d;
  int n_groups;
  long long grp_start[4];
  // BatchNorm statistics as a by-product of the forward epilogue: per (m-tile, lane quadrant) the column sums and sums of
  // squares of the STORED output values over the quadrant's real rows (padding rows of a group excluded), f32
  // [num_m_tiles * 4][2][n]. nullptr = off.
  float* bn_partials;
  long long grp_count[4];
};

template <bool BF16>
struct Cfg {
  static constexpr int kElem = BF16 ? 2 : 4;
  static constexpr int kSplit = BF16 ? 1 : 2;                // operand arrays (hi[, lo])
  static constexpr int BM = 128;
  static constexpr int BN = BF16 ? 256 : 128;
  static constexpr int BK = 128 / kElem;                     // one 128-byte swizzle row of K
  static constexpr int UK = 32 / kElem;                      // K per tcgen05.mma
  static constexpr int kChunk = 128 / kElem;                 // M/N elements per 128-byte line (MN-major)
  static constexpr int kABytes = BM * 128;                   // one operand tile (per hi/lo array)
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kSplit * (kABytes + kBBytes);
  static constexpr int kStages = BF16 ? 4 : 3;
  static constexpr int kTmemCols = 2 * BN;
  // The tensor core adds each MMA into the fp32 accumulator with truncation, so a long K chain drifts by
  // ~2^-24 per MMA (measured: 1e-4 at K=3584 with three TF32 passes). The fp32-grade mode therefore hands the
  // accumulator to the CUDA cores every kPromoteKBlocks k-blocks (round-to-nearest adds in registers).
  static constexpr int kPromoteKBlocks = BF16 ? (1 << 30) : 4;
  // epilogue staging: per epilogue warp one 32-row x 32-column chunk, rows padded by 16 B (conflict-free both ways)
  static constexpr int kStageRowF32 = 144, kStageRowBf16 = 80;
  static constexpr int kEpiStageBytes = 32 * kStageRowF32;     // per warp
  // warps 2-5 own full-size staging slots (fp32 or bf16 rows); warps 6-9 only ever stage bf16 rows
  static constexpr int kEpiStageBytesBf16 = 32 * kStageRowBf16;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + 4 * kEpiStageBytes +
                                    4 * kEpiStageBytesBf16;
};

// warp 0 TMA, warp 1 MMA, warps 2-5 epilogue, warps 6-9 second epilogue group (bf16 output of the bf16 kernel only:
// the two groups split the tile's columns — a short-K GEMM such as the data gradient (8 k-blocks per tile) is bound
// by how fast the accumulator leaves TMEM, not by the tensor pipe)
template <bool BF16> constexpr int kGemmThreads = BF16 ? 320 : 192;   // the tf32 kernel keeps its register budget

template <bool BF16>
__global__ void __launch_bounds__(kGemmThreads<BF16>, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg<BF16>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full = bars;                       // [kStages]
  uint64_t* empty = bars + C::kStages;         // [kStages]
  uint64_t* tmem_full = bars + 2 * C::kStages; // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::kSplit; ++s) { tma_prefetch_desc(&p.tm_a[s]); tma_prefetch_desc(&p.tm_b[s]); }
    for (int s = 0; s < C::kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    const uint32_t n_epi = (BF16 && p.out_bf16) ? 8 : 4;
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full + b, 1); mbar_init(tmem_empty + b, n_epi); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_mn = p.num_m_tiles * p.num_n_tiles;
  const int total_tiles = tiles_mn * p.num_splits;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int split = t / tiles_mn, mn = t - split * tiles_mn;
        const int m0 = (mn / p.num_n_tiles) * C::BM, n0 = (mn % p.num_n_tiles) * C::BN;
        const int kb0 = p.split_kb[split];
        const int kb1 = p.split_kb[split + 1];
        int grp = 0;
        if (p.remap_mode) {   // last group whose first padded row is <= m0 (empty groups share their successor's start)
          for (int g = p.n_groups - 1; g >= 0; --g)
            if (p.grp_start[g] <= m0) { grp = g; break; }
        }
        const int n0_b = p.remap_mode == 2 ? (n0 < p.remap_d ? grp * p.remap_d + n0 : n0 + 3 * p.remap_d) : n0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);
          uint8_t* st = smem + stage * C::kStageBytes;
          mbar_expect_tx(full + stage, C::kStageBytes);
          const int k0 = kb * C::BK;
          const int k0_b = p.remap_mode == 1 ? (k0 < p.remap_d ? grp * p.remap_d + k0 : k0 + 3 * p.remap_d) : k0;
#pragma unroll
          for (int s = 0; s < C::kSplit; ++s) {
            uint8_t* sa = st + s * C::kABytes;
            uint8_t* sb = st + C::kSplit * C::kABytes + s * C::kBBytes;
            if (p.a_mn_major) {
#pragma unroll
              for (int i = 0; i < C::BM / C::kChunk; ++i)   // one 128-byte-wide M chunk per box
                tma_load_2d(&p.tm_a[s], full + stage, sa + i * (C::BK * 128), m0 + i * C::kChunk, k0);
            } else {
              tma_load_2d(&p.tm_a[s], full + stage, sa, k0, m0);
            }
            if (p.b_mn_major) {
#pragma unroll
              for (int i = 0; i < C::BN / C::kChunk; ++i)
                tma_load_2d(&p.tm_b[s], full + stage, sb + i * (C::BK * 128), n0_b + i * C::kChunk, k0_b);
            } else {
              tma_load_2d(&p.tm_b[s], full + stage, sb, k0_b, n0_b);
            }
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      // descriptor geometry (bytes)
      const uint32_t a_lbo = p.a_mn_major ? C::BK * 128 : 16, b_lbo = p.b_mn_major ? C::BK * 128 : 16;
      const uint32_t a_kstep = p.a_mn_major ? C::UK * 128 : 32, b_kstep = p.b_mn_major ? C::UK * 128 : 32;
      int stage = 0, buf = 0;
      uint32_t phase = 0, buf_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int split = t / tiles_mn;
        const int kb0 = p.split_kb[split];
        const int kb1 = p.split_kb[split + 1];
        uint32_t d_tmem = 0, accum = 0;
        int in_chunk = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (in_chunk == 0) {  // start (a chunk of) an accumulation in a free TMEM buffer
            mbar_wait(tmem_empty + buf, buf_phase ^ 1);
            tc_fence_after();
            d_tmem = tmem_base + buf * C::BN;
            accum = 0;
          }
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + stage * C::kStageBytes);
#pragma unroll
          for (int k = 0; k < C::BK / C::UK; ++k) {
            const uint32_t a_hi = st + k * a_kstep;
            const uint32_t b_hi = st + C::kSplit * C::kABytes + k * b_kstep;
            const uint64_t da_hi = make_smem_desc(a_hi, a_lbo, 1024), db_hi = make_smem_desc(b_hi, b_lbo, 1024);
            umma<BF16>(d_tmem, da_hi, db_hi, p.idesc, accum);
            accum = 1;
            if constexpr (!BF16) {
              const uint64_t da_lo = make_smem_desc(a_hi + C::kABytes, a_lbo, 1024);
              const uint64_t db_lo = make_smem_desc(b_hi + C::kBBytes, b_lbo, 1024);
              umma<BF16>(d_tmem, da_hi, db_lo, p.idesc, 1);
              umma<BF16>(d_tmem, da_lo, db_hi, p.idesc, 1);
            }
          }
          umma_commit(empty + stage);  // smem stage reusable once these MMAs retire
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
          if (++in_chunk == C::kPromoteKBlocks || kb == kb1 - 1) {
            umma_commit(tmem_full + buf);  // (partial) accumulator complete
            buf ^= 1;
            if (buf == 0) buf_phase ^= 1;
            in_chunk = 0;
          }
        }
      }
    }
  } else if (warp < 6 || (BF16 && p.out_bf16)) {
    // ================================================================== epilogue (warps 2..5, and 6..9 for bf16 output)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const bool split_cols = BF16 && p.out_bf16;            // two warps per quadrant, half of the columns each
    const int col_lo = split_cols ? ((warp - 2) >> 2) * (C::BN / 2) : 0;
    const int col_hi = split_cols ? col_lo + C::BN / 2 : C::BN;
    int buf = 0;
    uint32_t buf_phase = 0;
    // One 32-row x 32-column chunk of the tile (thread = row after tcgen05.ld): bias / convert into a padded smem
    // staging buffer of this warp, then store with lanes running along the row — every global store instruction
    // writes whole 64/128-byte row segments instead of 32 scattered 16-byte pieces.
    uint8_t* stage = smem + C::kStages * C::kStageBytes + 256 +
                     (warp < 6 ? (warp - 2) * C::kEpiStageBytes : 4 * C::kEpiStageBytes + (warp - 6) * C::kEpiStageBytesBf16);
    // BatchNorm partials: column sums of the stored values over the real rows of this warp's quadrant, accumulated in
    // registers across the tiles this (persistent) CTA processes and flushed when the column block changes / at the end
    constexpr int kBnChunks = C::BN / 32;
    int bn_valid = 0;           // real rows of this warp's quadrant in the current tile
    int bn_n0 = -1;             // column block the accumulators belong to
    float bn_s1[kBnChunks], bn_s2[kBnChunks];
#pragma unroll
    for (int i = 0; i < kBnChunks; ++i) { bn_s1[i] = 0.f; bn_s2[i] = 0.f; }
    auto bn_flush = [&]() {
      if (bn_n0 < 0) return;
      float* dst = p.bn_partials + ((size_t)blockIdx.x * 4 + quad) * 2 * (size_t)p.n + bn_n0 + lane;
#pragma unroll
      for (int i = 0; i < kBnChunks; ++i) {
        if (bn_s1[i] != 0.f || bn_s2[i] != 0.f) {      // chunks outside this warp's column half stay untouched
          dst[32 * i] += bn_s1[i];                     // same thread, same address: plain read-modify-write
          dst[p.n + 32 * i] += bn_s2[i];
        }
        bn_s1[i] = 0.f;
        bn_s2[i] = 0.f;
      }
    };
    auto bn_chunk = [&](int ci, bool bf16_rows) {
      // column (chunk ci, lane) of the staged 32 x 32 chunk, summed over the quadrant's real rows in row order
      float s1 = 0.f, s2 = 0.f;
      for (int r = 0; r < bn_valid; ++r) {
        const float x = bf16_rows ? __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(stage + r * C::kStageRowBf16 + 2 * lane))
                                  : *reinterpret_cast<const float*>(stage + r * C::kStageRowF32 + 4 * lane);
        s1 += x;
        s2 += x * x;
      }
#pragma unroll
      for (int i = 0; i < kBnChunks; ++i)
        if (i == ci) { bn_s1[i] += s1; bn_s2[i] += s2; }
    };
    auto store_chunk = [&](const float* v, int64_t row0, int col, int split, int ci) {
      if (p.out_bf16) {
        uint8_t* mine = stage + lane * C::kStageRowBf16;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
          if (p.bias) { b0 = ldg4(p.bias + col + j); b1 = ldg4(p.bias + col + j + 4); }
          uint4 q;
          q.x = pack_bf16x2(v[j + 0] + b0.x, v[j + 1] + b0.y);
          q.y = pack_bf16x2(v[j + 2] + b0.z, v[j + 3] + b0.w);
          q.z = pack_bf16x2(v[j + 4] + b1.x, v[j + 5] + b1.y);
          q.w = pack_bf16x2(v[j + 6] + b1.z, v[j + 7] + b1.w);
          *reinterpret_cast<uint4*>(mine + 2 * j) = q;
        }
        __syncwarp();
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)split * p.split_stride + col;
#pragma unroll
        for (int i = 0; i < 4; ++i) {            // 8 rows x 64 B per instruction
          const int r = i * 8 + (lane >> 2), c16 = lane & 3;
          const uint4 q = *reinterpret_cast<const uint4*>(stage + r * C::kStageRowBf16 + c16 * 16);
          if (row0 + r < p.m) *reinterpret_cast<uint4*>(o + (size_t)(row0 + r) * p.ldd + c16 * 8) = q;
        }
        if (p.bn_partials) bn_chunk(ci, true);
      } else {
        uint8_t* mine = stage + lane * C::kStageRowF32;
        float4 bq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bq[j] = p.bias ? ldg4(p.bias + col + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(mine + 16 * j) = make_float4(v[4 * j] + bq[j].x, v[4 * j + 1] + bq[j].y,
                                                                   v[4 * j + 2] + bq[j].z, v[4 * j + 3] + bq[j].w);
        __syncwarp();
        float* o = reinterpret_cast<float*>(p.out) + (size_t)split * p.split_stride + col;
#pragma unroll
        for (int i = 0; i < 8; ++i) {            // 4 rows x 128 B per instruction
          const int r = i * 4 + (lane >> 3), c16 = lane & 7;
          const float4 q = *reinterpret_cast<const float4*>(stage + r * C::kStageRowF32 + c16 * 16);
          if (row0 + r < p.m) *reinterpret_cast<float4*>(o + (size_t)(row0 + r) * p.ldd + c16 * 4) = q;
        }
        if (p.bn_partials) bn_chunk(ci, false);
      }
      __syncwarp();   // staging buffer is reused by the next chunk
    };
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int split = t / tiles_mn, mn = t - split * tiles_mn;
      const int64_t m0 = (int64_t)(mn / p.num_n_tiles) * C::BM;
      const int n0 = (mn % p.num_n_tiles) * C::BN;
      const int kb0 = p.split_kb[split];
      const int kb1 = p.split_kb[split + 1];
      const int64_t row0 = m0 + quad * 32;      // first row of this warp's TMEM lane quadrant
      const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
      if (p.bn_partials) {
        long long lim = p.m;                     // end of the real rows this tile can hold
        if (p.n_groups > 0) {
          int grp = 0;
          for (int g = p.n_groups - 1; g >= 0; --g)
            if (p.grp_start[g] <= m0) { grp = g; break; }
          lim = p.grp_start[grp] + p.grp_count[grp];
        }
        const long long nv = lim - row0;
        bn_valid = nv < 0 ? 0 : (nv > 32 ? 32 : (int)nv);
        if (n0 != bn_n0) { bn_flush(); bn_n0 = n0; }
      }
      if constexpr (BF16) {
        // single accumulation: stream TMEM -> registers -> global
        mbar_wait(tmem_full + buf, buf_phase);
        tc_fence_after();
#pragma unroll
        for (int ci = 0; ci < kBnChunks; ++ci) {
          const int c0 = 32 * ci;
          if (c0 < col_lo || c0 >= col_hi) continue;            // warp-uniform: the other warp group's half
          uint32_t v[32];
          tmem_ld_32x32(lane_base + (uint32_t)(buf * C::BN + c0), v);
          if (row0 < p.m && n0 + c0 < p.n)   // warp-uniform; n is a multiple of 32 (host check): whole chunk in range
            store_chunk(reinterpret_cast<const float*>(v), row0, n0 + c0, split, ci);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty + buf);
        buf ^= 1;
        if (buf == 0) buf_phase ^= 1;
      } else {
        // promoted accumulation: add every TMEM chunk into fp32 registers (round-to-nearest)
        float acc[C::BN];
#pragma unroll
        for (int j = 0; j < C::BN; ++j) acc[j] = 0.f;
        const int n_chunks = (kb1 - kb0 + C::kPromoteKBlocks - 1) / C::kPromoteKBlocks;
        for (int ch = 0; ch < n_chunks; ++ch) {
          mbar_wait(tmem_full + buf, buf_phase);
          tc_fence_after();
#pragma unroll
          for (int c0 = 0; c0 < C::BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(lane_base + (uint32_t)(buf * C::BN + c0), v);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty + buf);
          buf ^= 1;
          if (buf == 0) buf_phase ^= 1;
        }
#pragma unroll
        for (int c0 = 0; c0 < C::BN; c0 += 32)
          if (row0 < p.m && n0 + c0 < p.n) store_chunk(acc + c0, row0, n0 + c0, split, c0 / 32);
      }
    }
    if (p.bn_partials) bn_flush();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// fixed-order reduction of split-K partials: out[i] = sum_s part[s][i]
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int n_splits, int64_t n_elems,
                                     float* __restrict__ out) {
  const int64_t n4 = n_elems >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = ldg4(part + 4 * i);
    for (int k = 1; k < n_splits; ++k) {
      const float4 v = ldg4(part + (size_t)k * n_elems + 4 * i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = s;
  }
}

// Structured weight gradient: partials [S][4d][d] from a split-K whose splits never straddle a row group.
//   rows [0, d)   (track block)  : summed per group g into out rows [g*d, (g+1)*d)
//   rows [d, 4d)  (onset, next, root blocks): summed over all splits into out rows [4d, 7d)
struct SplitGroups { int n_splits; int group[kMaxSplits]; };
__global__ void grouped_splitk_reduce_kernel(const float* __restrict__ part, const SplitGroups sg, int d,
                                             float* __restrict__ out) {
  const int n4 = d >> 2;
  const int64_t total = (int64_t)4 * d * n4;
  const size_t per_split = (size_t)4 * d * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / n4), c4 = (int)(i - (int64_t)row * n4);
    const float* src = part + (size_t)row * d + 4 * c4;
    if (row >= d) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < sg.n_splits; ++s) {
        const float4 v = ldg4(src + s * per_split);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      reinterpret_cast<float4*>(out + (size_t)(row + 3 * d) * d)[c4] = acc;
    } else {
      float4 acc[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < sg.n_splits; ++s) {
        const float4 v = ldg4(src + s * per_split);
        const int g = sg.group[s];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q == g) { acc[q].x += v.x; acc[q].y += v.y; acc[q].z += v.z; acc[q].w += v.w; }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) reinterpret_cast<float4*>(out + (size_t)(g * d + row) * d)[c4] = acc[g];
    }
  }
}

// dst[c][r] = src[r][c]  (fp32, 32x32 tiles through padded smem). Only the PB_F32 weight-gradient path uses
// it: tcgen05 kind::tf32 with MN-major operands needs a different shared-memory atom than the 16-byte-swizzled
// tiles used here, so that mode contracts over K-major transposed copies instead.
__global__ void __launch_bounds__(256) transpose_f32_kernel(const float* __restrict__ src, int64_t rows, int64_t cols,
                                                           int64_t ld_src, float* __restrict__ dst, int64_t ld_dst) {
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t r = r0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (r < rows && c < cols) ? src[(size_t)r * ld_src + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t c = c0 + ty + i, r = r0 + tx;
    if (c < cols && r < rows) dst[(size_t)c * ld_dst + r] = tile[tx][ty + i];
  }
}

static int transpose_f32(const float* src, int64_t rows, int64_t cols, int64_t ld_src, float* dst, int64_t ld_dst,
                         cudaStream_t st) {
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  PB_REQUIRE(grid.y <= 65535u * 32u, "transpose: too many rows");
  if (grid.y > 65535u) {  // fold very tall matrices over several launches
    const int64_t step = (int64_t)65535 * 32;
    for (int64_t r = 0; r < rows; r += step) {
      const int64_t nr = std::min(step, rows - r);
      dim3 g((unsigned)((cols + 31) / 32), (unsigned)((nr + 31) / 32));
      transpose_f32_kernel<<<g, 256, 0, st>>>(src + (size_t)r * ld_src, nr, cols, ld_src, dst + r, ld_dst);
      PB_LAUNCH_CHECK();
    }
    return PB_OK;
  }
  transpose_f32_kernel<<<grid, 256, 0, st>>>(src, rows, cols, ld_src, dst, ld_dst);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// Operand stored row-major as [rows, cols] with leading dimension ld (elements).
//   K-major use : rows = M/N index, cols = K          -> 2-D map {K, rows}, box {BK, box_rows}
//   MN-major use: rows = K index,  cols = M/N index   -> 2-D map {cols, K}, box {128-byte chunk, BK}; one box
//                                                        per chunk lands at chunk*BK*128 (the descriptor's LBO)
static int make_map(CUtensorMap* map, const void* ptr, bool bf16, bool mn_major, int64_t rows, int64_t cols, int64_t ld,
                    int box_mn, int bk) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (driver entry point lookup failed)");
    return PB_ERR_CUDA;
  }
  const int es = bf16 ? 2 : 4;
  const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r;
  if (!mn_major) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_mn};
    cuuint32_t estr[2] = {1, 1};
    r = fn(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const int chunk = 128 / es;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {(cuuint32_t)chunk, (cuuint32_t)bk};
    cuuint32_t estr[2] = {1, 1};
    r = fn(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld mn_major=%d)", (int)r,
              (long long)rows, (long long)cols, (long long)ld, (int)mn_major);
    return PB_ERR_CUDA;
  }
  return PB_OK;
}

struct Operand {
  const void* hi;
  const void* lo;
  int64_t rows, cols, ld;  // storage shape (row-major)
  bool mn_major;
};

template <bool BF16>
static int launch_gemm(const Operand& a, const Operand& b, int64_t m, int64_t n, int64_t k, void* out, int64_t ldd,
                       bool out_bf16, const float* bias, int num_splits, int64_t split_stride, cudaStream_t st,
                       const pb_groups_t* groups = nullptr, int remap_mode = 0, int remap_d = 0,
                       const int* split_table = nullptr, float* bn_partials = nullptr) {
  using C = Cfg<BF16>;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  for (int s = 0; s < C::kSplit; ++s) {
    rc = make_map(&p.tm_a[s], s ? a.lo : a.hi, BF16, a.mn_major, a.rows, a.cols, a.ld, C::BM, C::BK);
    if (rc) return rc;
    rc = make_map(&p.tm_b[s], s ? b.lo : b.hi, BF16, b.mn_major, b.rows, b.cols, b.ld, C::BN, C::BK);
    if (rc) return rc;
  }
  p.m = m;
  p.n = n;
  p.a_mn_major = a.mn_major;
  p.b_mn_major = b.mn_major;
  p.num_m_tiles = (int)((m + C::BM - 1) / C::BM);
  p.num_n_tiles = (int)((n + C::BN - 1) / C::BN);
  p.k_blocks = (int)((k + C::BK - 1) / C::BK);
  if (split_table) {   // caller-defined split boundaries (k-blocks), e.g. aligned to row groups
    p.num_splits = num_splits;
    for (int i = 0; i <= num_splits; ++i) p.split_kb[i] = split_table[i];
  } else {
    num_splits = std::max(1, std::min(std::min(num_splits, p.k_blocks), kMaxSplits));
    const int per = (p.k_blocks + num_splits - 1) / num_splits;
    p.num_splits = (p.k_blocks + per - 1) / per;
    for (int i = 0; i <= p.num_splits; ++i) p.split_kb[i] = std::min(i * per, p.k_blocks);
  }
  p.idesc = make_idesc(BF16, C::BM, C::BN, a.mn_major, b.mn_major);
  p.out = out;
  p.ldd = ldd;
  p.split_stride = split_stride;
  p.bias = bias;
  p.out_bf16 = out_bf16;
  if (groups && remap_mode) {
    p.remap_mode = remap_mode;
    p.remap_d = remap_d;
    p.n_groups = groups->n_groups;
    for (int g = 0; g < groups->n_groups && g < 4; ++g) { p.grp_start[g] = groups->start[g]; p.grp_count[g] = groups->count[g]; }
  }
  p.bn_partials = bn_partials;
  const int64_t total = (int64_t)p.num_m_tiles * p.num_n_tiles * p.num_splits;
  const int grid = (int)std::min<int64_t>(total, sm_count());
  // the dynamic shared-memory limit is a per-device function attribute: set it once on every device that launches
  static bool attr_set[64] = {};
  int dev = 0;
  PB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    PB_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  gemm_tcgen05_kernel<BF16><<<grid, kGemmThreads<BF16>, C::kSmemBytes, st>>>(p);
  PB_LAUNCH_CHECK();
  return p.num_splits;  // > 0
}

static int check_gemm_dims(int64_t m, int d, int k, int dtype, const char* who) {
  PB_REQUIRE(m > 0 && m < ((int64_t)1 << 31), "%s: m out of range", who);
  PB_REQUIRE(d >= 64 && d % 64 == 0, "%s: d=%d must be a multiple of 64", who, d);
  PB_REQUIRE(k >= 64 && k % 64 == 0, "%s: k=%d must be a multiple of 64", who, k);
  PB_REQUIRE(dtype == PB_BF16 || dtype == PB_F32, "%s: bad dtype", who);
  return PB_OK;
}

// structured layout: 4 groups, starts on tile boundaries, d a multiple of the widest tile so that no tile straddles
// two weight blocks, A width 4d
static int check_groups(const pb_groups_t* g, int64_t m, int d, int k, const char* who) {
  if (!g) return PB_OK;
  PB_REQUIRE(g->n_groups == 4, "%s: structured mode expects 4 row groups", who);
  PB_REQUIRE(d % 256 == 0, "%s: structured mode needs d %% 256 == 0 (got %d)", who, d);
  PB_REQUIRE(k == 4 * d, "%s: structured mode expects k == 4*d", who);
  for (int i = 0; i < 4; ++i)
    PB_REQUIRE(g->start[i] % 128 == 0 && g->start[i] >= 0 && g->start[i] + g->count[i] <= m, "%s: bad row group %d", who, i);
  return PB_OK;
}

static int bwd_weight_splits(int64_t m, int d, int k, bool bf16) {
  const int bm = 128, bn = bf16 ? 256 : 128, bk = bf16 ? 64 : 32;
  const int64_t tiles = (int64_t)((k + bm - 1) / bm) * ((d + bn - 1) / bn);
  const int64_t kblocks = (m + bk - 1) / bk;
  int64_t s = (3 * (int64_t)148 + tiles - 1) / tiles;       // ~3 waves of work items
  s = std::max<int64_t>(1, std::min<int64_t>(s, std::min<int64_t>(kblocks, 64)));
  return (int)s;
}

}  // namespace pb

using namespace pb;

static int rgcn_gemm_fwd_impl(const void* a_hi, const void* a_lo, int64_t lda, const void* wcat_t_hi,
                              const void* wcat_t_lo, const float* bias, void* out, int64_t ldo, int64_t m, int32_t d,
                              int32_t k, const pb_groups_t* groups, int32_t dtype, int32_t act_dtype, float* bn_partials,
                              pb_stream_t stream) {
  int rc = check_gemm_dims(m, d, k, dtype, "pb_rgcn_gemm_fwd");
  if (rc) return rc;
  PB_REQUIRE(act_dtype == PB_F32 || (act_dtype == PB_BF16 && dtype == PB_BF16),
             "pb_rgcn_gemm_fwd: act_dtype %d (a bf16 output needs the PB_BF16 operand mode)", act_dtype);
  const bool out_bf16 = act_dtype == PB_BF16;
  PB_REQUIRE(a_hi && wcat_t_hi && out, "pb_rgcn_gemm_fwd: null pointer");
  PB_REQUIRE(dtype == PB_BF16 || (a_lo && wcat_t_lo), "pb_rgcn_gemm_fwd: PB_F32 needs lo operands");
  PB_REQUIRE(lda >= k && lda % 8 == 0 && ldo >= d && ldo % 4 == 0, "pb_rgcn_gemm_fwd: bad leading dimension");
  int rc2 = check_groups(groups, m, d, k, "pb_rgcn_gemm_fwd");
  if (rc2) return rc2;
  const int kw = groups ? (groups->n_groups + 3) * d : k;     // K extent of the weight operand
  Operand a{a_hi, a_lo, m, k, lda, false};
  Operand b{wcat_t_hi, wcat_t_lo, d, kw, kw, false};
  cudaStream_t st = as_stream(stream);
  rc = dtype == PB_BF16 ? launch_gemm<true>(a, b, m, d, k, out, ldo, out_bf16, bias, 1, 0, st, groups, 1, d, nullptr, bn_partials)
                        : launch_gemm<false>(a, b, m, d, k, out, ldo, false, bias, 1, 0, st, groups, 1, d, nullptr, bn_partials);
  return rc < 0 ? rc : PB_OK;
}

extern "C" int pb_rgcn_gemm_fwd(const void* a_hi, const void* a_lo, int64_t lda, const void* wcat_t_hi,
                                const void* wcat_t_lo, const float* bias, void* out, int64_t ldo, int64_t m, int32_t d,
                                int32_t k, const pb_groups_t* groups, int32_t dtype, int32_t act_dtype,
                                pb_stream_t stream) {
  return rgcn_gemm_fwd_impl(a_hi, a_lo, lda, wcat_t_hi, wcat_t_lo, bias, out, ldo, m, d, k, groups, dtype, act_dtype, nullptr,
                            stream);
}

// one partial row per (persistent CTA, lane quadrant): the epilogue accumulates across the tiles a CTA processes
extern "C" int64_t pb_rgcn_gemm_fwd_bn_partial_rows(int64_t m) { return m <= 0 ? 0 : (int64_t)sm_count() * 4; }

extern "C" int pb_rgcn_gemm_fwd_bn(const void* a_hi, const void* a_lo, int64_t lda, const void* wcat_t_hi,
                                   const void* wcat_t_lo, const float* bias, void* out, int64_t ldo, int64_t m, int32_t d,
                                   int32_t k, const pb_groups_t* groups, int32_t dtype, int32_t act_dtype,
                                   float* bn_partials, pb_stream_t stream) {
  PB_REQUIRE(bn_partials, "pb_rgcn_gemm_fwd_bn: null bn_partials");
  return rgcn_gemm_fwd_impl(a_hi, a_lo, lda, wcat_t_hi, wcat_t_lo, bias, out, ldo, m, d, k, groups, dtype, act_dtype,
                            bn_partials, stream);
}

// Generic D[m,n] = A[m,k] . B[n,k]^T (+ bias[n]) on the same kernel: any nn.Linear forward / input gradient.
extern "C" int pb_gemm_nt(const void* a_hi, const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo,
                          int64_t ldb, const float* bias, void* out, int64_t ldo, int64_t m, int32_t n, int32_t k,
                          int32_t dtype, int32_t out_bf16, pb_stream_t stream) {
  int rc = check_gemm_dims(m, n, k, dtype, "pb_gemm_nt");
  if (rc) return rc;
  PB_REQUIRE(a_hi && b_hi && out, "pb_gemm_nt: null pointer");
  PB_REQUIRE(dtype == PB_BF16 || (a_lo && b_lo), "pb_gemm_nt: PB_F32 needs lo operands");
  PB_REQUIRE(lda >= k && lda % 8 == 0 && ldb >= k && ldb % 8 == 0 && ldo >= n && ldo % 8 == 0,
             "pb_gemm_nt: bad leading dimension");
  Operand a{a_hi, a_lo, m, k, lda, false};
  Operand b{b_hi, b_lo, n, k, ldb, false};
  cudaStream_t st = as_stream(stream);
  rc = dtype == PB_BF16 ? launch_gemm<true>(a, b, m, n, k, out, ldo, out_bf16 != 0, bias, 1, 0, st)
                        : launch_gemm<false>(a, b, m, n, k, out, ldo, out_bf16 != 0, bias, 1, 0, st);
  return rc < 0 ? rc : PB_OK;
}

extern "C" int pb_rgcn_gemm_bwd_data(const void* g_hi, const void* g_lo, int64_t ldg, const void* wcat_hi,
                                     const void* wcat_lo, void* d_a, int64_t ldda, int64_t m, int32_t d, int32_t k,
                                     const pb_groups_t* groups, int32_t dtype, pb_stream_t stream) {
  int rc = check_gemm_dims(m, d, k, dtype, "pb_rgcn_gemm_bwd_data");
  if (rc) return rc;
  PB_REQUIRE(g_hi && wcat_hi && d_a, "pb_rgcn_gemm_bwd_data: null pointer");
  PB_REQUIRE(dtype == PB_BF16 || (g_lo && wcat_lo), "pb_rgcn_gemm_bwd_data: PB_F32 needs lo operands");
  PB_REQUIRE(ldg >= d && ldg % 8 == 0 && ldda >= k && ldda % 8 == 0, "pb_rgcn_gemm_bwd_data: bad leading dimension");
  int rc2 = check_groups(groups, m, d, k, "pb_rgcn_gemm_bwd_data");
  if (rc2) return rc2;
  const int kw = groups ? (groups->n_groups + 3) * d : k;     // rows of the weight operand
  Operand a{g_hi, g_lo, m, d, ldg, false};          // [M, d], contraction over d
  Operand b{wcat_hi, wcat_lo, kw, d, d, false};     // Wcat [K, d]: rows = output columns, K-major in d
  cudaStream_t st = as_stream(stream);
  rc = dtype == PB_BF16 ? launch_gemm<true>(a, b, m, k, d, d_a, ldda, true, nullptr, 1, 0, st, groups, 2, d)
                        : launch_gemm<false>(a, b, m, k, d, d_a, ldda, false, nullptr, 1, 0, st, groups, 2, d);
  return rc < 0 ? rc : PB_OK;
}

static inline int64_t pad4(int64_t v) { return (v + 3) / 4 * 4; }

// split-K partials, plus (PB_F32 only) the K-major transposed operand copies
static size_t bwd_weight_ws(int64_t m, int32_t d, int32_t k, bool need_transposed) {
  const int s = std::max(bwd_weight_splits(m, d, k, true), bwd_weight_splits(m, d, k, false)) + 4;  // +4: per-group minimum
  const size_t partials = align_up((size_t)s * k * d * sizeof(float), 256);
  const size_t transposed = 2 * (align_up((size_t)k * pad4(m) * sizeof(float), 256) +
                                 align_up((size_t)d * pad4(m) * sizeof(float), 256));
  return partials + (need_transposed ? transposed : 0);
}

extern "C" size_t pb_rgcn_gemm_bwd_weight_workspace_bytes(int64_t m, int32_t d, int32_t k) {
  if (m <= 0 || d <= 0 || k <= 0) return 0;
  return bwd_weight_ws(m, d, k, true);
}

extern "C" size_t pb_rgcn_gemm_bwd_weight_workspace_bytes_for(int64_t m, int32_t d, int32_t k, int32_t dtype) {
  if (m <= 0 || d <= 0 || k <= 0) return 0;
  return bwd_weight_ws(m, d, k, dtype != PB_BF16);
}

extern "C" int pb_rgcn_gemm_bwd_weight(const void* a_hi, const void* a_lo, int64_t lda, const void* g_hi,
                                       const void* g_lo, int64_t ldg, float* d_wcat, int64_t m, int32_t d, int32_t k,
                                       const pb_groups_t* groups, int32_t dtype, void* workspace,
                                       size_t workspace_bytes, pb_stream_t stream) {
  int rc = check_gemm_dims(m, d, k, dtype, "pb_rgcn_gemm_bwd_weight");
  if (rc) return rc;
  if ((rc = check_groups(groups, m, d, k, "pb_rgcn_gemm_bwd_weight"))) return rc;
  PB_REQUIRE(a_hi && g_hi && d_wcat && workspace, "pb_rgcn_gemm_bwd_weight: null pointer");
  PB_REQUIRE(dtype == PB_BF16 || (a_lo && g_lo), "pb_rgcn_gemm_bwd_weight: PB_F32 needs lo operands");
  PB_REQUIRE(lda >= k && lda % 8 == 0 && ldg >= d && ldg % 8 == 0, "pb_rgcn_gemm_bwd_weight: bad leading dimension");
  PB_REQUIRE(workspace_bytes >= pb_rgcn_gemm_bwd_weight_workspace_bytes_for(m, d, k, dtype), "pb_rgcn_gemm_bwd_weight: workspace too small");
  const bool bf16 = dtype == PB_BF16;
  int splits = bwd_weight_splits(m, d, k, bf16);
  cudaStream_t st = as_stream(stream);
  float* part = reinterpret_cast<float*>(workspace);
  const int64_t n_elems = (int64_t)k * d;
  // structured layout: split boundaries follow the row groups so that the track block can be reduced per group
  int table[kMaxSplits + 1];
  SplitGroups sg;
  memset(&sg, 0, sizeof(sg));
  const int* split_table = nullptr;
  if (groups) {
    const int bk = bf16 ? 64 : 32;
    const int64_t n_valid = groups->count[0] + groups->count[1] + groups->count[2] + groups->count[3];
    const int budget = std::max(1, std::min(splits, kMaxSplits - 4));
    int ns = 0;
    for (int g = 0; g < 4; ++g) {
      if (groups->count[g] <= 0) continue;
      const int kb_begin = (int)(groups->start[g] / bk);   // group starts are multiples of 128
      const int kb_end = (int)((groups->start[g] + groups->count[g] + bk - 1) / bk);
      int cnt = (int)std::max<int64_t>(1, (int64_t)budget * groups->count[g] / std::max<int64_t>(1, n_valid));
      cnt = std::min(cnt, kb_end - kb_begin);
      for (int i = 0; i < cnt; ++i) {
        table[ns] = kb_begin + (int)((int64_t)(kb_end - kb_begin) * i / cnt);
        sg.group[ns++] = g;
      }
    }
    PB_REQUIRE(ns > 0, "pb_rgcn_gemm_bwd_weight: empty row groups");
    // split s covers [table[s], table[s+1]): the last split of a group runs on over the group's zero padding
    table[ns] = (int)((m + bk - 1) / bk);
    sg.n_splits = ns;
    splits = ns;
    split_table = table;
  }
  if (bf16) {
    Operand a{a_hi, a_lo, m, k, lda, true};   // stored [nodes, K]: output rows (K) contiguous -> MN-major
    Operand b{g_hi, g_lo, m, d, ldg, true};   // stored [nodes, d]
    rc = launch_gemm<true>(a, b, k, d, m, part, d, false, nullptr, splits, n_elems, st, nullptr, 0, 0, split_table);
  } else {
    // K-major transposed copies: At [K, m], gt [d, m]
    const int64_t mp = pad4(m);
    const int s_max = std::max(bwd_weight_splits(m, d, k, true), bwd_weight_splits(m, d, k, false)) + 4;
    char* wsp = reinterpret_cast<char*>(workspace) + align_up((size_t)s_max * k * d * sizeof(float), 256);
    float* at_hi = reinterpret_cast<float*>(wsp); wsp += align_up((size_t)k * mp * sizeof(float), 256);
    float* at_lo = reinterpret_cast<float*>(wsp); wsp += align_up((size_t)k * mp * sizeof(float), 256);
    float* gt_hi = reinterpret_cast<float*>(wsp); wsp += align_up((size_t)d * mp * sizeof(float), 256);
    float* gt_lo = reinterpret_cast<float*>(wsp);
    if ((rc = transpose_f32(reinterpret_cast<const float*>(a_hi), m, k, lda, at_hi, mp, st))) return rc;
    if ((rc = transpose_f32(reinterpret_cast<const float*>(a_lo), m, k, lda, at_lo, mp, st))) return rc;
    if ((rc = transpose_f32(reinterpret_cast<const float*>(g_hi), m, d, ldg, gt_hi, mp, st))) return rc;
    if ((rc = transpose_f32(reinterpret_cast<const float*>(g_lo), m, d, ldg, gt_lo, mp, st))) return rc;
    Operand a{at_hi, at_lo, k, m, mp, false};
    Operand b{gt_hi, gt_lo, d, m, mp, false};
    rc = launch_gemm<false>(a, b, k, d, m, part, d, false, nullptr, splits, n_elems, st, nullptr, 0, 0, split_table);
  }
  if (rc < 0) return rc;
  const int used = rc;
  const unsigned grid = (unsigned)std::min<int64_t>((n_elems / 4 + 255) / 256, (int64_t)sm_count() * 8);
  if (groups)
    grouped_splitk_reduce_kernel<<<grid, 256, 0, st>>>(part, sg, d, d_wcat);   // d_wcat is [7d, d] here
  else
    splitk_reduce_kernel<<<grid, 256, 0, st>>>(part, used, n_elems, d_wcat);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
