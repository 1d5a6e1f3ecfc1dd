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
          mbar_wait_backoff(empty + st, (uint32_t)(((sidx / kTcStages) & 1) ^ 1));
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
#define PB_TC(DR, CPL, AB) launch_tc<DR, CPL, AB>(g, x, table, d_a, ldda, gy_res, gx, partials, bits, scale, st)
#define PB_TC_D(DR, AB) (d == 512 ? PB_TC(DR, 4, AB) : PB_TC(DR, 2, AB))
  if (act_bf16) return bits ? PB_TC_D(true, true) : PB_TC_D(false, true);
  return bits ? PB_TC_D(true, false) : PB_TC_D(false, false);
#undef PB_TC_D
#undef PB_TC
}

}  // namespace pb
