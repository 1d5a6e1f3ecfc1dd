// Aggregation backward with the edge-table gradient accumulated on the tensor cores (sm_100a, bf16 mode).
//
// Backward of GCL.message + scatter-mean (reference model.py:123-135, propagate model.py:110; formulas SURVEY.md §3.4):
//   gx[u]  = gy_res[u] + dA[u, root block] + sum_{e: src_e = u} ds_e * T[dist_e]
//   dT[k]  = sum_{e: dist_e = k} q_e,   q_e = ds_e * x[src_e],   ds_e = dH[dst_e, slot_e] / |segment| * keep_e * 1[x*T > 0]
//
// The first line is a gather by source (warp per source, lanes over channels, exactly the arithmetic of the legacy
// agg_bwd_dx_kernel). The second is a 32-bin reduction of E rows of d channels. The legacy path writes the rows q_e to
// HBM and reads them back grouped by distance (2 x E x d bytes of pure overhead, half of the kernel's traffic). Here
// the reduction is a matrix product,  dT^T [d, 32] = Q^T [d, E] . onehot(dist) [E, 32],  and it runs on the otherwise
// idle tensor pipe: the gather warps drop each q_e (bf16, as before) into a shared-memory stage laid out as an
// MN-major UMMA operand (128-byte swizzle) next to a one-hot line of its distance; one elected thread issues
// tcgen05.mma per full stage (M = 128 channels x N = 64 (32 distances, padded to the 128-byte line) x K = 16 edges),
// and the [d, 32] fp32 accumulator never leaves TMEM until the CTA is done. No atomics: an edge's stage slot is its
// position in the CTA's (visit-ordered) edge range, so the summation order is fixed -> bit-reproducible.
//
// One persistent CTA per SM: 14 gather warps (128 registers each, no spills: with the 214 KB shared-memory carve-out L1 is
// ~14 KB and every spilled word is an L2 round trip — 16 / 24 warps measure 400 / 460 us against 340) + 1 MMA warp + 1
// L2-prefetch warp. Decomposition measured on the LMD16 batch-256 graph (d = 512): gather + math alone 316 us, staging +
// barriers + MMA +22 us, removing the row loads entirely changes nothing: the kernel is bound by the ~300 instructions a
// warp spends per edge (16 channels per lane) at ~0.5 IPC per scheduler, not by HBM. Shared memory: the edge table T [32, d] fp32 (read per
// edge by every lane), 4 stages x 32 edges of Q (4 x 32 KB at d = 512) and their one-hot lines (4 x 4 KB).
#include "common.cuh"
#include "mbar.cuh"
#include "tc.cuh"

namespace pb {

#ifndef PB_TC_SLEEP
#define PB_TC_SLEEP 256
#endif
#ifndef PB_TC_WARPS
#define PB_TC_WARPS 14
#endif
constexpr int kTcProducers = PB_TC_WARPS;              // gather warps per CTA
constexpr int kTcThreads = (kTcProducers + 2) * 32;    // + the MMA warp + the L2 prefetch warp
#ifndef PB_TC_AHEAD
#define PB_TC_AHEAD 64
#endif
constexpr int kTcAhead = PB_TC_AHEAD;                           // sources the prefetch warp may run ahead of the finished ones
#ifndef PB_TC_STAGES
#define PB_TC_STAGES 4
#define PB_TC_STAGE_EDGES 32
#endif
constexpr int kTcStages = PB_TC_STAGES;
constexpr int kTcStageEdges = PB_TC_STAGE_EDGES;       // K extent of a stage: k-steps of 16
constexpr int kTcN = 64;                               // 32 distances padded to one 128-byte MN-major line
constexpr int kTcBStage = kTcStageEdges * 128;         // one-hot lines of a stage (bytes)

__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t a) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
// wait that backs off: a gather warp whose stage is still in use must not compete for issue slots with the warps
// that are filling it
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  while (!ok) {
    __nanosleep(PB_TC_SLEEP);
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
// pull `bytes` (multiple of 16, 16-byte aligned) at p into L2; no destination, no completion tracking
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// CPL = float4 chunks per lane, d == 128 * CPL. ABF: x / gy_res / gx stored as bf16 (else fp32). d_a is bf16.
template <bool DROPOUT, int CPL, bool ABF>
__global__ void __launch_bounds__(kTcThreads, 1) agg_bwd_tc_kernel(
    const int4* __restrict__ visit_meta, const int* __restrict__ visit_edge_ptr, const int4* __restrict__ out_rec,
    const void* __restrict__ x, const float* __restrict__ table, const __nv_bfloat16* __restrict__ d_a, int64_t ldda,
    const void* __restrict__ gy_res, void* __restrict__ gx, float* __restrict__ partials, int64_t n_nodes, int n_rel,
    const uint16_t* __restrict__ keep_bits, float keep_scale, uint32_t idesc) {
  constexpr uint32_t kFull = 0xffffffffu;
  constexpr int d = 128 * CPL;
  constexpr int G16 = (CPL + 3) / 4;
  constexpr int kMB = d / 128;                          // M-blocks (accumulators) of 128 channels
  constexpr int kChunkBytes = kTcStageEdges * 128;      // one 64-channel chunk of a stage: [32 edges][128 B]
  constexpr int kQStage = (d / 64) * kChunkBytes;       // bytes of Q per stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_st = smem;                                            // [kTcStages][kQStage]
  uint8_t* b_st = q_st + kTcStages * kQStage;                      // [kTcStages][kTcBStage]
  float* t_s = reinterpret_cast<float*>(b_st + kTcStages * kTcBStage);   // [32][d]
  // sign masks of the table in the lanes' channel order: word [dist][lane] = (bits of T > 0) | (bits of T < 0) << 16,
  // bit 4 j + i <-> channel 4 (lane + 32 j) + i  (the order of the lane's keep-bit word)
  uint32_t* t_sign = reinterpret_cast<uint32_t*>(t_s + PB_N_DISTS * d);  // [32][32]
  uint2* mask_lut = reinterpret_cast<uint2*>(t_sign + PB_N_DISTS * 32);   // [16]: 4 keep bits -> bf16-pair AND masks
  uint64_t* bars = reinterpret_cast<uint64_t*>(mask_lut + 16);
  uint64_t* full = bars;                       // [kTcStages], one arrival per edge slot
  uint64_t* empty = bars + kTcStages;          // [kTcStages], tcgen05.commit
  uint64_t* done = bars + 2 * kTcStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  volatile int* progress = reinterpret_cast<volatile int*>(tmem_slot + 1);   // sources finished by the gather warps
  // stages whose MMAs have been issued. A gather warp may be more than a lap of the stage ring ahead of the tensor core
  // (skewed out-degrees: one warp works through a 400-edge hub while the others are done with everything behind it), and the
  // parity of `empty` alone cannot tell "the previous use is done" from "two uses ago is done": the warp first waits until
  // the previous use of its stage has at least been issued, then the parity wait is unambiguous.
  volatile int* mma_front = reinterpret_cast<volatile int*>(tmem_slot + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t per = (n_nodes + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per;
  const int64_t r1 = r0 + per < n_nodes ? r0 + per : n_nodes;
  int e0 = 0, n_edges = 0;
  if (r0 < r1) {
    e0 = __ldg(visit_edge_ptr + r0);
    n_edges = __ldg(visit_edge_ptr + r1) - e0;
  }
  const int n_total = (n_edges + kTcStageEdges - 1) / kTcStageEdges;   // stages this CTA fills

  if (warp == kTcProducers && lane == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(full + s, kTcStageEdges); mbar_init(empty + s, 1); }
    mbar_init(done, 1);
    *progress = 0;
    *mma_front = 0;
    fence_barrier_init();
  }
  if (warp == kTcProducers) tmem_alloc(tmem_slot, kMB * kTcN < 32 ? 32 : kMB * kTcN);
  // edge table -> shared memory; stages start out zeroed (an unwritten slot must hold finite values)
  for (int i = threadIdx.x; i < PB_N_DISTS * d / 4; i += kTcThreads)
    reinterpret_cast<float4*>(t_s)[i] = __ldg(reinterpret_cast<const float4*>(table) + i);
  for (int i = threadIdx.x; i < kTcStages * (kQStage + kTcBStage) / 16; i += kTcThreads)
    reinterpret_cast<uint4*>(q_st)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < PB_N_DISTS * 32; i += kTcThreads) {
    const int k = i >> 5, l = i & 31;
    uint32_t pos = 0, neg = 0;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(table + (size_t)k * d + 4 * (l + 32 * j)));
      pos |= ((t.x > 0.f) | ((t.y > 0.f) << 1) | ((t.z > 0.f) << 2) | ((t.w > 0.f) << 3)) << (4 * j);
      neg |= ((t.x < 0.f) | ((t.y < 0.f) << 1) | ((t.z < 0.f) << 2) | ((t.w < 0.f) << 3)) << (4 * j);
    }
    t_sign[i] = pos | (neg << 16);
  }
  if (threadIdx.x < 16) {
    const uint32_t b = threadIdx.x;
    mask_lut[b] = make_uint2(((b & 1u) ? 0xFFFFu : 0u) | ((b & 2u) ? 0xFFFF0000u : 0u),
                             ((b & 4u) ? 0xFFFFu : 0u) | ((b & 8u) ? 0xFFFF0000u : 0u));
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t q_u32 = smem_u32(q_st), b_u32 = smem_u32(b_st);

  if (warp == kTcProducers) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      for (int g = 0; g < n_total; ++g) {
        const int st = g & (kTcStages - 1);
        mbar_wait(full + st, (uint32_t)((g / kTcStages) & 1));
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < kTcStageEdges / 16; ++ks) {
          const uint64_t b_desc = make_smem_desc(b_u32 + st * kTcBStage + ks * 2048, kTcBStage, 1024);
#pragma unroll
          for (int mb = 0; mb < kMB; ++mb) {
            const uint64_t a_desc = make_smem_desc(q_u32 + st * kQStage + 2 * mb * kChunkBytes + ks * 2048, kChunkBytes, 1024);
            umma<true>(tmem_base + mb * kTcN, a_desc, b_desc, idesc, (g > 0 || ks > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty + st);        // the stage may be refilled once these MMAs have read it
        *mma_front = g + 1;
      }
      if (n_total > 0) umma_commit(done);
    }
  } else if (warp == kTcProducers + 1) {
    // ================================================================== L2 prefetch warp
    // The gather warps are latency-bound: every source is a chain meta -> records -> rows of dependent loads. This
    // warp walks the same work list a bounded distance ahead (lane = source) and pulls every row the gather warps
    // will touch into L2 with bulk prefetches, so that their loads are L2 hits instead of DRAM round trips.
    constexpr uint32_t kRowA = d * (ABF ? 2 : 4), kRowG = d * 2;
    for (int64_t b0 = r0; b0 < r1; b0 += 32) {
      while ((int)(b0 - r0) >= *progress + kTcAhead) __nanosleep(200);
      const int64_t vi = b0 + lane;
      if (vi < r1) {
        const int4 m = __ldg(visit_meta + vi);
        prefetch_l2_bulk(static_cast<const char*>(x) + (size_t)m.x * kRowA, kRowA);
        if (gy_res) prefetch_l2_bulk(static_cast<const char*>(gy_res) + (size_t)m.x * kRowA, kRowA);
        prefetch_l2_bulk(d_a + (size_t)m.x * ldda + (size_t)n_rel * d, kRowG);
        // all records of the source first (independent loads, one round trip), then the prefetches they address
        int4 r[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) r[s] = s < m.z ? __ldg(out_rec + m.y + s) : make_int4(0, 0, 0, 0);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s < m.z) {
            prefetch_l2_bulk(d_a + (size_t)r[s].x * ldda + (size_t)(r[s].y & 0xff) * d, kRowG);
            if constexpr (DROPOUT) prefetch_l2_bulk(keep_bits + (size_t)(uint32_t)r[s].z * G16 * 32, G16 * 64);
          }
        }
        for (int s = 8; s < m.z; ++s) {
          const int4 q = __ldg(out_rec + m.y + s);
          prefetch_l2_bulk(d_a + (size_t)q.x * ldda + (size_t)(q.y & 0xff) * d, kRowG);
          if constexpr (DROPOUT) prefetch_l2_bulk(keep_bits + (size_t)(uint32_t)q.z * G16 * 32, G16 * 64);
        }
      }
    }
  } else {
    // ================================================================== gather warps (one source at a time)
    const int last_pos = n_edges - 1;
    int stage_ok = -1;                                             // last stage (p / 32) whose buffer is known to be free
    // per-lane constants of the Q-row layout: channel chunk c4 = lane + 32 j lives in 64-channel block c4 / 16, 16-byte
    // unit (c4 % 16) / 2 of the edge's 128-byte line (units XOR-swizzled with the line index), half (c4 & 1)
    uint32_t q_off[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int c4 = lane + 32 * j;
      q_off[j] = (uint32_t)(c4 >> 4) * kChunkBytes + (uint32_t)(((c4 & 15) >> 1) << 4) + (uint32_t)(c4 & 1) * 8u;
    }
    const uint32_t b_off = (uint32_t)((lane >> 2) << 4) + (uint32_t)(lane & 3) * 4u;
    struct EdgeRegs { uint2 raw[CPL]; uint32_t kw; int meta, cnt; };
    // The dependent chain of a source is  {row, first edge, degree} -> edge records -> rows.  Its first two links are
    // taken off the critical path by running them ahead: the metadata of the source after next and the records of the
    // next source are loaded while the current one is processed, so a source starts by issuing ALL its first loads
    // (features, residual, root block, the rows of its first three edges) at once — one memory round trip instead of
    // three.
    auto load_meta = [&](int64_t vi, int4& m, int& vep) {
      if (vi < r1) { m = __ldg(visit_meta + vi); vep = __ldg(visit_edge_ptr + vi); } else { m = make_int4(0, 0, 0, 0); vep = 0; }
    };
    auto load_rec = [&](const int4& m) {
      return lane < m.z ? __ldg(out_rec + m.y + lane) : make_int4(0, 0, 0, 0);
    };
    int4 m_cur, m_nxt, m_nn, rec_cur, rec_nxt;
    int vep_cur, vep_nxt, vep_nn;
    load_meta(r0 + warp, m_cur, vep_cur);
    load_meta(r0 + warp + kTcProducers, m_nxt, vep_nxt);
    rec_cur = load_rec(m_cur);
    for (int64_t vi = r0 + warp; vi < r1; vi += kTcProducers) {
      load_meta(vi + 2 * kTcProducers, m_nn, vep_nn);             // two sources ahead
      rec_nxt = load_rec(m_nxt);                                  // one source ahead (its metadata arrived a source ago)
      const int u = m_cur.x, beg = m_cur.y, end = m_cur.y + m_cur.z;
      const int pos0 = vep_cur - e0 - beg;                        // position of out-edge i in the CTA's range: pos0 + i
      int base = beg;
      int4 my_rec = rec_cur;
      float4 xu[CPL], acc[CPL];
      // 1[x * T > 0] = (x > 0 and T > 0) or (x < 0 and T < 0): sign bits of this source's features, once per source
      uint32_t xpos = 0, xneg = 0;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const size_t c4 = 4 * (size_t)(lane + 32 * j);
        xu[j] = act_ld4<ABF>(x, (size_t)u * d + c4);
        const uint2 rp = __ldg(reinterpret_cast<const uint2*>(d_a + (size_t)u * ldda + (size_t)n_rel * d + c4));   // root branch
        const float2 ra = unpack_bf16x2(rp.x), rb = unpack_bf16x2(rp.y);
        acc[j] = make_float4(ra.x, ra.y, rb.x, rb.y);
        if (gy_res) {
          const float4 r = act_ld4_stream<ABF>(gy_res, (size_t)u * d + c4);                                         // residual branch
          acc[j].x += r.x; acc[j].y += r.y; acc[j].z += r.z; acc[j].w += r.w;
        }
        // integer view: v > 0 <=> bits in (0, 0x7f800000]; v < 0 <=> bits as unsigned > 0x80000000 (NaN never occurs)
        const int bx = __float_as_int(xu[j].x), by = __float_as_int(xu[j].y), bz = __float_as_int(xu[j].z), bw = __float_as_int(xu[j].w);
        xpos |= (uint32_t)((bx > 0) | ((by > 0) << 1) | ((bz > 0) << 2) | ((bw > 0) << 3)) << (4 * j);
        xneg |= (uint32_t)(((uint32_t)bx > 0x80000000u) | (((uint32_t)by > 0x80000000u) << 1) | (((uint32_t)bz > 0x80000000u) << 2) |
                           (((uint32_t)bw > 0x80000000u) << 3)) << (4 * j);
      }
      // software pipeline over the out-edges: the gradient row and keep-bits of edge i+1 are in flight while edge i
      // is consumed (two register sets, used alternately: no copies)
      auto fetch = [&](EdgeRegs& e, int i) {
        if (i - base == 32) {  // warp-uniform (out-degree > 32 only)
          base = i;
          my_rec = base + lane < end ? __ldg(out_rec + base + lane) : make_int4(0, 0, 0, 0);
        }
        const int dst = __shfl_sync(kFull, my_rec.x, i - base);
        e.meta = __shfl_sync(kFull, my_rec.y, i - base);
        e.cnt = __shfl_sync(kFull, my_rec.w, i - base);
        const __nv_bfloat16* grow = d_a + (size_t)dst * ldda + (size_t)(e.meta & 0xff) * d + 4 * lane;
#pragma unroll
        for (int j = 0; j < CPL; ++j) e.raw[j] = __ldg(reinterpret_cast<const uint2*>(grow + 128 * j));
        if constexpr (DROPOUT) {
          const uint32_t eid = (uint32_t)__shfl_sync(kFull, my_rec.z, i - base);
          e.kw = __ldg(keep_bits + (size_t)eid * G16 * 32 + lane);
        }
      };
      int pend_n = 0;                                             // slots written but not yet published (one stage)
      auto process = [&](const EdgeRegs& e, int i) {
        const int dist = (e.meta >> 8) & (PB_N_DISTS - 1);
        const float* trow = t_s + dist * d + 4 * lane;
        float coef = e.cnt > 1 ? 1.0f / (float)e.cnt : 1.0f;      // d(mean)/d(sum), one reciprocal per edge
        if constexpr (DROPOUT) coef *= keep_scale;
        // stage slot of this edge: its position in the CTA's edge range
        const int p = pos0 + i;
        const int sidx = p / kTcStageEdges, st = sidx & (kTcStages - 1), slot = p & (kTcStageEdges - 1);
        if (sidx != stage_ok) {                                    // warp-uniform: first slot this warp writes in the stage
          while (*mma_front <= sidx - kTcStages) __nanosleep(PB_TC_SLEEP);    // its previous use has been issued ...
          mbar_wait_backoff(empty + st, (uint32_t)(((sidx / kTcStages) & 1) ^ 1));   // ... and has been read
          stage_ok = sidx;
        }
        const uint32_t swz = (uint32_t)(slot & 7) << 4;
        const uint32_t q_row = q_u32 + st * kQStage + slot * 128;
        // keep decision of the lane's 16 channels as one bit mask (dropout bit and 1[x * T > 0]), applied to the packed
        // bf16 gradient words before they are unpacked: a dropped channel is an exact +0 in everything that follows
        const uint32_t ts = t_sign[dist * 32 + lane];
        uint32_t keep = (xpos & ts) | (xneg & (ts >> 16));
        if constexpr (DROPOUT) keep &= e.kw;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const uint2 mk = mask_lut[(keep >> (4 * j)) & 15u];       // 4 keep bits -> two words of 0xFFFF fields
          const float2 da = unpack_bf16x2(e.raw[j].x & mk.x), db = unpack_bf16x2(e.raw[j].y & mk.y);
          const float4 ds = make_float4(da.x * coef, da.y * coef, db.x * coef, db.y * coef);
          const float4 t = *reinterpret_cast<const float4*>(trow + 128 * j);
          const float4 xv = xu[j];
          acc[j].x += ds.x * t.x; acc[j].y += ds.y * t.y; acc[j].z += ds.z * t.z; acc[j].w += ds.w * t.w;
          sts64(q_row + (q_off[j] ^ swz), pack_bf16x2(ds.x * xv.x, ds.y * xv.y), pack_bf16x2(ds.z * xv.z, ds.w * xv.w));
        }
        // one-hot line of the distance: elements n = 2 lane, 2 lane + 1 (lanes >= 16 would write n >= 32: always zero)
        if (lane < 16) {
          const uint32_t v = (dist == 2 * lane ? 0x3F80u : 0u) | (dist == 2 * lane + 1 ? 0x3F800000u : 0u);
          sts32(b_u32 + st * kTcBStage + slot * 128 + (b_off ^ swz), v);
        }
        // publish: the proxy fence orders this warp's generic-proxy stores against the tensor core's async-proxy reads;
        // the slots a source fills in one stage are published together
        ++pend_n;
        const bool tail = p == last_pos && slot != kTcStageEdges - 1;         // the CTA's last, partial stage
        if (i + 1 == end || slot == kTcStageEdges - 1 || tail) {              // warp-uniform
          if (tail) {
            for (int k = slot + 1; k < kTcStageEdges; ++k)
              if (lane < 16) sts32(b_u32 + st * kTcBStage + k * 128 + (b_off ^ ((uint32_t)(k & 7) << 4)), 0u);
            pend_n += kTcStageEdges - 1 - slot;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive_n(full + st, (uint32_t)pend_n);
          pend_n = 0;
        }
      };
      EdgeRegs ea, eb, ec;                                        // three rows in flight
      if (beg < end) fetch(ea, beg);
      if (beg + 1 < end) fetch(eb, beg + 1);
      for (int i = beg; i < end; i += 3) {
        if (i + 2 < end) fetch(ec, i + 2);
        process(ea, i);
        if (i + 1 >= end) break;
        if (i + 3 < end) fetch(ea, i + 3);
        process(eb, i + 1);
        if (i + 2 >= end) break;
        if (i + 4 < end) fetch(eb, i + 4);
        process(ec, i + 2);
      }
#pragma unroll
      for (int j = 0; j < CPL; ++j) act_st4_stream<ABF>(gx, (size_t)u * d + 4 * (lane + 32 * j), acc[j]);
      if (lane == 0) atomicAdd(const_cast<int*>(progress), 1);
      m_cur = m_nxt; vep_cur = vep_nxt; rec_cur = rec_nxt;
      m_nxt = m_nn; vep_nxt = vep_nn;
    }
    // ================================================================== epilogue: TMEM -> partials [cta][32][d]
    if (warp < 4) {
      if (n_total > 0) {
        mbar_wait(done, 0);
        tc_fence_after();
      }
      float* out = partials + (size_t)blockIdx.x * PB_N_DISTS * d;
#pragma unroll 1
      for (int mb = 0; mb < kMB; ++mb) {
        uint32_t v[32];
        if (n_total > 0) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)(mb * kTcN), v);
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = 0u;
        }
        const int c = mb * 128 + 32 * warp + lane;      // TMEM lane = channel inside the M-block, column = distance
#pragma unroll
        for (int k = 0; k < PB_N_DISTS; ++k) out[(size_t)k * d + c] = __uint_as_float(v[k]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kTcProducers) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kMB * kTcN < 32 ? 32 : kMB * kTcN);
  }
}

// ====================================================================================================================
// Ring variant (bf16 activations): the gradient rows reach the SM without passing through registers.
//
// In the kernel above a gather warp has at most three 1 KB rows in flight (its registers), ~40 KB per SM: the warps
// sit on the long scoreboard (3.1 stalled warps per issue, ncu) while HBM idles at 34 %. Here every gather warp copies
// the gradient rows dH[dst, slot] of the source it will process NEXT with cp.async (16 bytes per lane, no registers,
// no L1) straight into the Q stage slots those edges own — already in the swizzled MN-major position of their
// channels — together with their keep bits, and then processes the current source: it reads its 16 channels of an
// edge from shared memory and writes q = ds * x back IN PLACE (same address per lane). The four Q stages (128 edges,
// 128 KB) are landing buffer and UMMA operand at once. The records of a source (row, distance, 1 / |segment|, edge id)
// come from the plan's record stream (pb_csr_bwd_stream) two sources ahead, its three per-source rows (features,
// root block, residual) are loaded into registers one source ahead.
//
// Flow control needs no extra barriers: a warp waits for its own copies (cp.async groups), and before copying into
// a stage it checks that the stage's previous MMAs are done (`empty`). That check is a non-blocking test when the warp
// still has a source to process (then the copies are simply issued after it) and a blocking wait otherwise — a
// blocked warp only ever waits for positions older than everything it still has to consume, so the oldest unconsumed
// edge can always proceed. Sources with more than 32 out-edges are processed in chunks of 32 (copy, wait, consume).
// The parity of `empty` is only trusted once the shared counter `mma_front` says the stage's previous use has been
// issued (a warp can be more than a lap of the ring ahead of the tensor core when out-degrees are skewed).
#ifndef PB_RG_WARPS
#define PB_RG_WARPS 15
#endif
constexpr int kRgWarps = PB_RG_WARPS;                  // gather warps
constexpr int kRgThreads = (kRgWarps + 1) * 32;        // + the MMA warp
constexpr int kRgSlots = kTcStages * kTcStageEdges;    // edges resident in the Q stages

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
// read-only tables (written once before the CTA-wide barrier): not volatile, the compiler may schedule them freely
__device__ __forceinline__ uint32_t lds32_ro(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64_ro(uint32_t addr) {
  uint2 v;
  asm("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds128f_ro(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

template <bool DROPOUT, int CPL>
__global__ void __launch_bounds__(kRgThreads, 1) agg_bwd_ring_kernel(
    const int4* __restrict__ visit_meta, const int* __restrict__ visit_edge_ptr, const int4* __restrict__ stream,
    const __nv_bfloat16* __restrict__ x, const float* __restrict__ table, const __nv_bfloat16* __restrict__ d_a,
    const __nv_bfloat16* __restrict__ gy_res, __nv_bfloat16* __restrict__ gx, float* __restrict__ partials, int64_t n_nodes,
    int n_rel, const uint16_t* __restrict__ keep_bits, float keep_scale, uint32_t idesc) {
  constexpr uint32_t kFull = 0xffffffffu;
  constexpr int d = 128 * CPL;
  constexpr int G16 = (CPL + 3) / 4;
  constexpr int kMB = d / 128;
  constexpr int kChunkBytes = kTcStageEdges * 128;
  constexpr int kQStage = (d / 64) * kChunkBytes;
  constexpr int kBitsSlot = G16 * 64;                   // keep bits of an edge (bytes)
  constexpr uint32_t kRow = d * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_st = smem;                                            // [kTcStages][kQStage]
  uint8_t* b_st = q_st + kTcStages * kQStage;                      // [kTcStages][kTcBStage]
  float* t_s = reinterpret_cast<float*>(b_st + kTcStages * kTcBStage);   // [32][d]
  uint32_t* t_sign = reinterpret_cast<uint32_t*>(t_s + PB_N_DISTS * d);  // [32][32], as above
  uint2* mask_lut = reinterpret_cast<uint2*>(t_sign + PB_N_DISTS * 32);   // [16]
  uint8_t* bits_ring = reinterpret_cast<uint8_t*>(mask_lut + 16);         // [kRgSlots][kBitsSlot]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bits_ring + kRgSlots * kBitsSlot);
  uint64_t* full = bars;                       // [kTcStages]: q written, one arrival per edge slot
  uint64_t* empty = bars + kTcStages;          // [kTcStages]: tcgen05.commit -> the stage may be refilled
  uint64_t* done = bars + 2 * kTcStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  volatile int* mma_front = reinterpret_cast<volatile int*>(tmem_slot + 1);   // stages whose MMAs have been issued (see above)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t per = (n_nodes + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per;
  const int64_t r1 = r0 + per < n_nodes ? r0 + per : n_nodes;
  int e0 = 0, n_edges = 0;
  if (r0 < r1) {
    e0 = __ldg(visit_edge_ptr + r0);
    n_edges = __ldg(visit_edge_ptr + r1) - e0;
  }
  const int n_total = (n_edges + kTcStageEdges - 1) / kTcStageEdges;

  if (warp == kRgWarps && lane == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(full + s, kTcStageEdges); mbar_init(empty + s, 1); }
    mbar_init(done, 1);
    *mma_front = 0;
    fence_barrier_init();
  }
  if (warp == kRgWarps) tmem_alloc(tmem_slot, kMB * kTcN < 32 ? 32 : kMB * kTcN);
  for (int i = threadIdx.x; i < PB_N_DISTS * d / 4; i += kRgThreads)
    reinterpret_cast<float4*>(t_s)[i] = __ldg(reinterpret_cast<const float4*>(table) + i);
  for (int i = threadIdx.x; i < kTcStages * (kQStage + kTcBStage) / 16; i += kRgThreads)
    reinterpret_cast<uint4*>(q_st)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < PB_N_DISTS * 32; i += kRgThreads) {
    const int k = i >> 5, l = i & 31;
    uint32_t pos = 0, neg = 0;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(table + (size_t)k * d + 4 * (l + 32 * j)));
      pos |= ((t.x > 0.f) | ((t.y > 0.f) << 1) | ((t.z > 0.f) << 2) | ((t.w > 0.f) << 3)) << (4 * j);
      neg |= ((t.x < 0.f) | ((t.y < 0.f) << 1) | ((t.z < 0.f) << 2) | ((t.w < 0.f) << 3)) << (4 * j);
    }
    t_sign[i] = pos | (neg << 16);
  }
  if (threadIdx.x < 16) {
    const uint32_t b = threadIdx.x;
    mask_lut[b] = make_uint2(((b & 1u) ? 0xFFFFu : 0u) | ((b & 2u) ? 0xFFFF0000u : 0u),
                             ((b & 4u) ? 0xFFFFu : 0u) | ((b & 8u) ? 0xFFFF0000u : 0u));
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t q_u32 = smem_u32(q_st), b_u32 = smem_u32(b_st);
  const uint32_t bits_u32 = smem_u32(bits_ring);
  const uint32_t trow_u32 = smem_u32(t_s) + 16 * lane, sign_u32 = smem_u32(t_sign) + 4 * lane, lut_u32 = smem_u32(mask_lut);

  if (warp == kRgWarps) {
    // ================================================================== MMA issuer (as above)
    if (lane == 0) {
      for (int g = 0; g < n_total; ++g) {
        const int st = g & (kTcStages - 1);
#ifdef PB_RG_MMA_SPIN
        mbar_wait(full + st, (uint32_t)((g / kTcStages) & 1));
#else
        mbar_wait_backoff(full + st, (uint32_t)((g / kTcStages) & 1));
#endif
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < kTcStageEdges / 16; ++ks) {
          const uint64_t b_desc = make_smem_desc(b_u32 + st * kTcBStage + ks * 2048, kTcBStage, 1024);
#pragma unroll
          for (int mb = 0; mb < kMB; ++mb) {
            const uint64_t a_desc = make_smem_desc(q_u32 + st * kQStage + 2 * mb * kChunkBytes + ks * 2048, kChunkBytes, 1024);
            umma<true>(tmem_base + mb * kTcN, a_desc, b_desc, idesc, (g > 0 || ks > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty + st);
        *mma_front = g + 1;
      }
      if (n_total > 0) umma_commit(done);
    }
  } else {
    // ================================================================== gather warps (one source at a time)
    const int last_pos = n_edges - 1;
    uint32_t q_off[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int c4 = lane + 32 * j;
      q_off[j] = (uint32_t)(c4 >> 4) * kChunkBytes + (uint32_t)(((c4 & 15) >> 1) << 4) + (uint32_t)(c4 & 1) * 8u;
    }
    // copy: lane moves the 16-byte units lane, lane + 32 (, ...) of a row = channels 8 k .. 8 k + 7 -> 64-channel block
    // k / 8, unit k % 8 of the edge's line; unit k + 32 sits four blocks further
    const uint32_t cp_off = (uint32_t)(lane >> 3) * kChunkBytes + (uint32_t)((lane & 7) << 4);
    const uint32_t b_off = (uint32_t)((lane >> 2) << 4) + (uint32_t)(lane & 3) * 4u;
    struct Hdr { uint2 x[CPL], root[CPL], res[CPL]; };
    struct Meta { int u, deg, vep; };
    auto load_meta = [&](int64_t vi) {
      Meta m;
      if (vi < r1) {
        const int4 t = __ldg(visit_meta + vi);
        m.u = t.x; m.deg = t.z; m.vep = __ldg(visit_edge_ptr + vi);
      } else { m.u = 0; m.deg = 0; m.vep = 0; }
      return m;
    };
    // records of edges c0 .. c0 + 31 of source vi (lane = edge): {row block of dH, kind | dist << 8, edge id, 1/|segment|}
    auto load_rec = [&](int64_t vi, const Meta& m, int c0) {
      return c0 + lane < m.deg ? __ldg(stream + 3 * vi + m.vep + 3 + c0 + lane) : make_int4(0, 0, 0, 0);
    };
    auto load_hdr = [&](Hdr& h, int u) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const size_t c4 = 4 * (size_t)(lane + 32 * j);
        h.x[j] = __ldg(reinterpret_cast<const uint2*>(x + (size_t)u * d + c4));
        h.root[j] = __ldg(reinterpret_cast<const uint2*>(d_a + ((size_t)u * (n_rel + 1) + n_rel) * d + c4));
        h.res[j] = gy_res ? ld_stream2(gy_res + (size_t)u * d + c4) : make_uint2(0u, 0u);
      }
    };
    int load_ok = -1;                                              // last stage index this warp knows to be free
    // copies rows + keep bits of edges [0, n) of `rec` (n <= 32) into the slots of positions pbase ..; one cp.async group.
    // Non-blocking mode gives up (false, nothing issued) when a stage still holds edges the tensor core has not read.
    auto issue = [&](const int4& rec, int n, int pbase, bool blocking) -> bool {
      if (n > 0) {
        const int s_last = (pbase + n - 1) / kTcStageEdges;
        for (int sx = max(pbase / kTcStageEdges, load_ok + 1); sx <= s_last; ++sx) {
          uint64_t* bar = empty + (sx & (kTcStages - 1));
          const uint32_t par = (uint32_t)(((sx / kTcStages) & 1) ^ 1);
          // the stage's previous use must have been issued before the parity of `empty` means "it has been read"
          if (blocking) {
            while (*mma_front <= sx - kTcStages) __nanosleep(PB_TC_SLEEP);
            mbar_wait_backoff(bar, par);
          } else if (*mma_front <= sx - kTcStages || !mbar_test(bar, par)) {
            return false;
          }
          load_ok = sx;
        }
        for (int i = 0; i < n; ++i) {
          const uint32_t rx = (uint32_t)__shfl_sync(kFull, rec.x, i);
          const int p = pbase + i;
          const int st = (p / kTcStageEdges) & (kTcStages - 1), slot = p & (kTcStageEdges - 1);
          const uint32_t dst = q_u32 + st * kQStage + slot * 128 + (cp_off ^ ((uint32_t)(slot & 7) << 4));
          const char* src = reinterpret_cast<const char*>(d_a) + (size_t)rx * kRow + 16 * lane;
#pragma unroll
          for (int h = 0; h < d / 256; ++h) cp_async16(dst + h * 4 * kChunkBytes, src + 512 * h);
        }
        if constexpr (DROPOUT) {
          if (lane < n) {
            const uint32_t dst = bits_u32 + (uint32_t)((pbase + lane) & (kRgSlots - 1)) * kBitsSlot;
            const char* src = reinterpret_cast<const char*>(keep_bits) + (size_t)(uint32_t)rec.z * kBitsSlot;
#pragma unroll
            for (int q = 0; q < kBitsSlot / 16; ++q) cp_async16(dst + 16 * q, src + 16 * q);
          }
        }
      }
      cp_async_commit();
      return true;
    };

    const int64_t v0 = r0 + warp;
    Meta m0 = load_meta(v0), m1 = load_meta(v0 + kRgWarps), m2 = load_meta(v0 + 2 * kRgWarps), m3;
    int4 rec_cur = load_rec(v0, m0, 0), rec_nxt = load_rec(v0 + kRgWarps, m1, 0), rec_nn;
    Hdr hdr;
    load_hdr(hdr, m0.u);
    issue(rec_cur, min(m0.deg, 32), m0.vep - e0, true);
    for (int64_t vi = v0; vi < r1; vi += kRgWarps) {
      m3 = load_meta(vi + 3 * kRgWarps);
      rec_nn = load_rec(vi + 2 * kRgWarps, m2, 0);
      const int u = m0.u, deg = m0.deg;
      const int pos0 = m0.vep - e0;
      float4 xu[CPL], acc[CPL];
      uint32_t xpos = 0, xneg = 0;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const float2 xa = unpack_bf16x2(hdr.x[j].x), xb = unpack_bf16x2(hdr.x[j].y);
        xu[j] = make_float4(xa.x, xa.y, xb.x, xb.y);
        const float2 ra = unpack_bf16x2(hdr.root[j].x), rb = unpack_bf16x2(hdr.root[j].y);
        const float2 sa = unpack_bf16x2(hdr.res[j].x), sb2 = unpack_bf16x2(hdr.res[j].y);
        acc[j] = make_float4(ra.x + sa.x, ra.y + sa.y, rb.x + sb2.x, rb.y + sb2.y);
        const int bx = __float_as_int(xu[j].x), by = __float_as_int(xu[j].y), bz = __float_as_int(xu[j].z), bw = __float_as_int(xu[j].w);
        xpos |= (uint32_t)((bx > 0) | ((by > 0) << 1) | ((bz > 0) << 2) | ((bw > 0) << 3)) << (4 * j);
        xneg |= (uint32_t)(((uint32_t)bx > 0x80000000u) | (((uint32_t)by > 0x80000000u) << 1) | (((uint32_t)bz > 0x80000000u) << 2) |
                           (((uint32_t)bw > 0x80000000u) << 3)) << (4 * j);
      }
      load_hdr(hdr, m1.u);                                         // next source's rows: in flight during this one
      // next source's gradient rows: start them now if their stages are free, else after this source
      const bool eager = issue(rec_nxt, min(m1.deg, 32), m1.vep - e0, false);
      if (eager) cp_async_wait<1>(); else cp_async_wait<0>();      // this source's copies (every lane: its own part)
      __syncwarp();
      int pend_n = 0;
      for (int c0 = 0; c0 < deg; c0 += 32) {
        const int n = min(deg - c0, 32);
        if (c0 > 0) {                                              // out-degree > 32: synchronous chunks
          rec_cur = load_rec(vi, m0, c0);
          issue(rec_cur, n, pos0 + c0, true);
          cp_async_wait<0>();
          __syncwarp();
        }
        for (int i = 0; i < n; ++i) {
          const int p = pos0 + c0 + i;
          const int st = (p / kTcStageEdges) & (kTcStages - 1), slot = p & (kTcStageEdges - 1);
          const int dist = (__shfl_sync(kFull, rec_cur.y, i) >> 8) & (PB_N_DISTS - 1);
          float coef = __int_as_float(__shfl_sync(kFull, rec_cur.w, i));
          if constexpr (DROPOUT) coef *= keep_scale;
          const uint32_t swz = (uint32_t)(slot & 7) << 4;
          const uint32_t q_row = q_u32 + st * kQStage + slot * 128;
          uint2 raw[CPL];
#pragma unroll
          for (int j = 0; j < CPL; ++j) raw[j] = lds64(q_row + (q_off[j] ^ swz));
          const uint32_t trow = trow_u32 + dist * (d * 4);
          const uint32_t ts = lds32_ro(sign_u32 + dist * 128);
          uint32_t keep = (xpos & ts) | (xneg & (ts >> 16));
          if constexpr (DROPOUT) keep &= lds16(bits_u32 + (uint32_t)(p & (kRgSlots - 1)) * kBitsSlot + 2 * lane);
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const uint2 mk = lds64_ro(lut_u32 + (j == 0 ? (keep << 3) & 0x78u : (keep >> (4 * j - 3)) & 0x78u));
            const float2 da = unpack_bf16x2(raw[j].x & mk.x), db = unpack_bf16x2(raw[j].y & mk.y);
            const float4 ds = make_float4(da.x * coef, da.y * coef, db.x * coef, db.y * coef);
            const float4 t = lds128f_ro(trow + 512 * j);
            const float4 xv = xu[j];
            acc[j].x += ds.x * t.x; acc[j].y += ds.y * t.y; acc[j].z += ds.z * t.z; acc[j].w += ds.w * t.w;
            sts64(q_row + (q_off[j] ^ swz), pack_bf16x2(ds.x * xv.x, ds.y * xv.y), pack_bf16x2(ds.z * xv.z, ds.w * xv.w));
          }
          if (lane < 16) {
            const uint32_t v = (dist == 2 * lane ? 0x3F80u : 0u) | (dist == 2 * lane + 1 ? 0x3F800000u : 0u);
            sts32(b_u32 + st * kTcBStage + slot * 128 + (b_off ^ swz), v);
          }
          ++pend_n;
          const bool tail = p == last_pos && slot != kTcStageEdges - 1;
          if (c0 + i + 1 == deg || slot == kTcStageEdges - 1 || tail) {       // warp-uniform
            if (tail) {
              for (int k = slot + 1; k < kTcStageEdges; ++k)
                if (lane < 16) sts32(b_u32 + st * kTcBStage + k * 128 + (b_off ^ ((uint32_t)(k & 7) << 4)), 0u);
              pend_n += kTcStageEdges - 1 - slot;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_n(full + st, (uint32_t)pend_n);
            pend_n = 0;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < CPL; ++j)
        st_stream2(gx + (size_t)u * d + 4 * (lane + 32 * j),
                   make_uint2(pack_bf16x2(acc[j].x, acc[j].y), pack_bf16x2(acc[j].z, acc[j].w)));
      if (!eager) issue(rec_nxt, min(m1.deg, 32), m1.vep - e0, true);
      m0 = m1; m1 = m2; m2 = m3;
      rec_cur = rec_nxt; rec_nxt = rec_nn;
    }
    cp_async_wait<0>();
    // ================================================================== epilogue: TMEM -> partials [cta][32][d]
    if (warp < 4) {
      if (n_total > 0) {
        mbar_wait(done, 0);
        tc_fence_after();
      }
      float* out = partials + (size_t)blockIdx.x * PB_N_DISTS * d;
#pragma unroll 1
      for (int mb = 0; mb < kMB; ++mb) {
        uint32_t v[32];
        if (n_total > 0) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)(mb * kTcN), v);
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = 0u;
        }
        const int c = mb * 128 + 32 * warp + lane;
#pragma unroll
        for (int k = 0; k < PB_N_DISTS; ++k) out[(size_t)k * d + c] = __uint_as_float(v[k]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kRgWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kMB * kTcN < 32 ? 32 : kMB * kTcN);
  }
}

static size_t rg_smem_bytes(int d) {
  const int g16 = (d / 128 + 3) / 4;
  return (size_t)kTcStages * ((size_t)(d / 64) * kTcStageEdges * 128 + kTcBStage) + (size_t)PB_N_DISTS * d * sizeof(float) +
         PB_N_DISTS * 32 * sizeof(uint32_t) + 128 /*mask LUT*/ + (size_t)kRgSlots * (g16 * 64) /*keep bits*/ +
         256 /*barriers, TMEM slot*/ + 1024 /*alignment*/;
}

template <bool DROP, int CPL>
static int launch_ring(const pb_csr_t* g, const void* x, const float* table, const void* d_a, const void* gy_res, void* gx,
                       float* partials, const uint16_t* bits, float scale, cudaStream_t st) {
  const int d = 128 * CPL;
  const size_t smem = rg_smem_bytes(d);
  static bool attr_set[64] = {};
  int dev = 0;
  PB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    PB_CUDA(cudaFuncSetAttribute(agg_bwd_ring_kernel<DROP, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const uint32_t idesc = make_idesc(true, 128, kTcN, 1, 1);
  agg_bwd_ring_kernel<DROP, CPL><<<agg_bwd_tc_ctas(g->n_nodes), kRgThreads, smem, st>>>(
      reinterpret_cast<const int4*>(g->visit_meta), g->visit_edge_ptr, reinterpret_cast<const int4*>(g->bwd_stream),
      static_cast<const __nv_bfloat16*>(x), table, static_cast<const __nv_bfloat16*>(d_a),
      static_cast<const __nv_bfloat16*>(gy_res), static_cast<__nv_bfloat16*>(gx), partials, g->n_nodes, g->n_relations, bits,
      scale, idesc);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

// PB200_AGG_BWD_RING=0 keeps the register-gather kernel (A/B measurements)
static bool ring_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PB200_AGG_BWD_RING");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static size_t tc_smem_bytes(int d) {
  return (size_t)kTcStages * ((size_t)(d / 64) * kTcStageEdges * 128 + kTcBStage) + (size_t)PB_N_DISTS * d * sizeof(float) +
         PB_N_DISTS * 32 * sizeof(uint32_t) /*sign masks*/ + 128 /*mask LUT*/ + 256 /*barriers*/ + 1024 /*alignment*/;
}

int agg_bwd_tc_ctas(int64_t n_nodes) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n_nodes + kTcProducers - 1) / kTcProducers, (int64_t)sm_count()));
}

bool agg_bwd_tc_eligible(int d, int dtype) { return dtype == PB_BF16 && (d == 256 || d == 512); }

template <bool DROP, int CPL, bool ABF>
static int launch_tc(const pb_csr_t* g, const void* x, const float* table, const void* d_a, int64_t ldda, const void* gy_res,
                     void* gx, float* partials, const uint16_t* bits, float scale, cudaStream_t st) {
  const int d = 128 * CPL;
  const size_t smem = tc_smem_bytes(d);
  static bool attr_set[64] = {};
  int dev = 0;
  PB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    PB_CUDA(cudaFuncSetAttribute(agg_bwd_tc_kernel<DROP, CPL, ABF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const uint32_t idesc = make_idesc(true, 128, kTcN, 1, 1);
  agg_bwd_tc_kernel<DROP, CPL, ABF><<<agg_bwd_tc_ctas(g->n_nodes), kTcThreads, smem, st>>>(
      reinterpret_cast<const int4*>(g->visit_meta), g->visit_edge_ptr, reinterpret_cast<const int4*>(g->out_rec), x, table,
      reinterpret_cast<const __nv_bfloat16*>(d_a), ldda, gy_res, gx, partials, g->n_nodes, g->n_relations, bits, scale, idesc);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

// called by pb_agg_bwd_fused (aggregate.cu) when agg_bwd_tc_eligible()
int agg_bwd_tc_launch(const pb_csr_t* g, const void* x, int d, const float* table, const void* d_a, int64_t ldda,
                      const void* gy_res, void* gx, float* partials, const uint16_t* bits, float scale, bool act_bf16,
                      cudaStream_t st) {
  if (act_bf16 && g->bwd_stream && ldda == (int64_t)(g->n_relations + 1) * d && ring_enabled()) {
    if (d == 512) return bits ? launch_ring<true, 4>(g, x, table, d_a, gy_res, gx, partials, bits, scale, st)
                              : launch_ring<false, 4>(g, x, table, d_a, gy_res, gx, partials, bits, scale, st);
    return bits ? launch_ring<true, 2>(g, x, table, d_a, gy_res, gx, partials, bits, scale, st)
                : launch_ring<false, 2>(g, x, table, d_a, gy_res, gx, partials, bits, scale, st);
  }
#define PB_TC(DR, CPL, AB) launch_tc<DR, CPL, AB>(g, x, table, d_a, ldda, gy_res, gx, partials, bits, scale, st)
#define PB_TC_D(DR, AB) (d == 512 ? PB_TC(DR, 4, AB) : PB_TC(DR, 2, AB))
  if (act_bf16) return bits ? PB_TC_D(true, true) : PB_TC_D(false, true);
  return bits ? PB_TC_D(true, false) : PB_TC_D(false, false);
#undef PB_TC_D
#undef PB_TC
}

}  // namespace pb
