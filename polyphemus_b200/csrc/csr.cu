// CSR plan from an edge list: replaces the per-relation boolean compaction of the reference
// (masked_edge_index / masked_edge_attrs, model.py:30-38,104-105 — one nonzero() host sync per relation
// per layer) and PyG's index_select / scatter bookkeeping inside propagate (model.py:110).
//
// Two views of the same edges, both deterministic (edges of a segment ordered by their edge_index column):
//   destination view: segments keyed (dst, relation)  -> forward mean aggregation without atomics
//   source view:      segments keyed src              -> backward scatter-by-source without atomics
// Integer atomics are used only to hand out slots; a per-segment sort by edge id removes the order
// nondeterminism before anything floating-point sees the data.
#include "common.cuh"
#include "scan.cuh"

namespace pb {

__device__ __forceinline__ bool edge_ok(long long s, long long d, int t, int dist, int64_t n, int r) {
  return s >= 0 && s < n && d >= 0 && d < n && t < r && dist < PB_N_DISTS;
}

__global__ void csr_count_kernel(const long long* __restrict__ src, const long long* __restrict__ dst,
                                 const uint8_t* __restrict__ type, const uint8_t* __restrict__ dist,
                                 int64_t n_nodes, int64_t n_edges, int n_rel, int* __restrict__ seg_cnt,
                                 int* __restrict__ out_cnt) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    const long long s = src[e], d = dst[e];
    const int t = type[e];
    if (!edge_ok(s, d, t, dist[e], n_nodes, n_rel)) continue;
    atomicAdd(seg_cnt + d * n_rel + t, 1);
    atomicAdd(out_cnt + s, 1);
  }
}

__global__ void csr_slot_kernel(const long long* __restrict__ src, const long long* __restrict__ dst,
                                const uint8_t* __restrict__ type, const uint8_t* __restrict__ dist,
                                int64_t n_nodes, int64_t n_edges, int n_rel, const int* __restrict__ in_ptr,
                                const int* __restrict__ out_ptr, int* __restrict__ in_cur, int* __restrict__ out_cur,
                                int* __restrict__ in_eid, int* __restrict__ out_eid) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    const long long s = src[e], d = dst[e];
    const int t = type[e];
    if (!edge_ok(s, d, t, dist[e], n_nodes, n_rel)) continue;
    const long long key = d * n_rel + t;
    in_eid[in_ptr[key] + atomicAdd(in_cur + key, 1)] = (int)e;
    out_eid[out_ptr[s] + atomicAdd(out_cur + s, 1)] = (int)e;
  }
}

// in-place ascending sort of a (usually tiny) segment: insertion sort, heap sort for long segments
__device__ void sort_segment(int* a, int n) {
  if (n <= 32) {
    for (int i = 1; i < n; ++i) {
      int v = a[i], j = i - 1;
      while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; }
      a[j + 1] = v;
    }
    return;
  }
  auto sift = [&](int start, int end) {
    int root = start;
    while (2 * root + 1 <= end) {
      int child = 2 * root + 1, sw = root;
      if (a[sw] < a[child]) sw = child;
      if (child + 1 <= end && a[sw] < a[child + 1]) sw = child + 1;
      if (sw == root) return;
      int tmp = a[root]; a[root] = a[sw]; a[sw] = tmp;
      root = sw;
    }
  };
  for (int start = (n - 2) / 2; start >= 0; --start) sift(start, n - 1);
  for (int end = n - 1; end > 0; --end) {
    int tmp = a[end]; a[end] = a[0]; a[0] = tmp;
    sift(0, end - 1);
  }
}

__global__ void csr_finish_in_kernel(const long long* __restrict__ src, const uint8_t* __restrict__ dist,
                                     int64_t n_segments, const int* __restrict__ in_ptr, int* __restrict__ in_eid,
                                     int* __restrict__ in_edge) {
  for (int64_t seg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; seg < n_segments;
       seg += (int64_t)gridDim.x * blockDim.x) {
    const int beg = in_ptr[seg], end = in_ptr[seg + 1];
    if (end == beg) continue;
    sort_segment(in_eid + beg, end - beg);
    for (int i = beg; i < end; ++i) {
      const int e = in_eid[i];
      in_edge[i] = (int)src[e] | ((int)dist[e] << 26);
    }
  }
}

__global__ void csr_finish_out_kernel(const long long* __restrict__ dst, const uint8_t* __restrict__ type,
                                      const uint8_t* __restrict__ dist, int64_t n_nodes, int n_rel,
                                      const int* __restrict__ in_ptr, const int* __restrict__ out_ptr,
                                      int* __restrict__ out_eid, int4* __restrict__ out_rec) {
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n_nodes;
       u += (int64_t)gridDim.x * blockDim.x) {
    const int beg = out_ptr[u], end = out_ptr[u + 1];
    if (end == beg) continue;
    sort_segment(out_eid + beg, end - beg);
    for (int i = beg; i < end; ++i) {
      const int e = out_eid[i];
      const long long d = dst[e];
      const int t = type[e];
      const long long key = d * n_rel + t;
      out_rec[i] = make_int4((int)d, t | ((int)dist[e] << 8), e, in_ptr[key + 1] - in_ptr[key]);
    }
  }
}

// ---- distance grouping of the source view (for the deterministic edge-table gradient) -----------------------
// The backward writes one row per out-edge position; the table gradient dT[k] is the sum of the rows whose edge
// has distance k. dist_perm lists the positions grouped by distance (stable), dist_item splits the groups into
// kDistItems contiguous work items of similar size, item_ptr[k] indexes the items of distance k.
constexpr int kDistBlock = 1024;

__global__ void __launch_bounds__(kDistBlock) dist_hist_kernel(const int4* __restrict__ out_rec, int64_t n_edges,
                                                               int* __restrict__ block_hist) {
  __shared__ int hist[PB_N_DISTS];
  if (threadIdx.x < PB_N_DISTS) hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * kDistBlock + threadIdx.x;
  if (p < n_edges) atomicAdd(hist + ((out_rec[p].y >> 8) & 31), 1);
  __syncthreads();
  if (threadIdx.x < PB_N_DISTS) block_hist[(size_t)blockIdx.x * PB_N_DISTS + threadIdx.x] = hist[threadIdx.x];
}

// one warp per distance: exclusive scan of its column of block_hist (in place), group sizes -> dist_ptr, items
__global__ void __launch_bounds__(1024) dist_offsets_kernel(int* __restrict__ block_hist, int64_t n_blocks,
                                                            int64_t n_edges, int n_items, int* __restrict__ dist_ptr,
                                                            int* __restrict__ item_ptr, int4* __restrict__ items) {
  __shared__ int total[PB_N_DISTS];
  __shared__ int start[PB_N_DISTS + 1];
  __shared__ int n_it[PB_N_DISTS + 1];
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int carry = 0;
  for (int64_t b0 = 0; b0 < n_blocks; b0 += 32) {
    const int64_t b = b0 + lane;
    const int v = b < n_blocks ? block_hist[(size_t)b * PB_N_DISTS + k] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (b < n_blocks) block_hist[(size_t)b * PB_N_DISTS + k] = carry + incl - v;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) total[k] = carry;
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0, it = 0;
    const int spare = n_items - PB_N_DISTS;   // every non-empty group gets >= 1 item, the rest go by size
    for (int j = 0; j < PB_N_DISTS; ++j) {
      start[j] = acc;
      n_it[j] = it;
      if (total[j] > 0) {
        int extra = (int)((int64_t)spare * total[j] / (n_edges > 0 ? n_edges : 1));
        int cnt = 1 + extra;
        if (cnt > total[j]) cnt = total[j];
        it += cnt;
      }
      acc += total[j];
    }
    start[PB_N_DISTS] = acc;
    n_it[PB_N_DISTS] = it;
  }
  __syncthreads();
  if (threadIdx.x <= PB_N_DISTS) {
    dist_ptr[threadIdx.x] = start[threadIdx.x];
    item_ptr[threadIdx.x] = n_it[threadIdx.x];
  }
  // items of distance k: even split of [start[k], start[k+1])
  const int cnt = n_it[k + 1] - n_it[k];
  const int64_t len = start[k + 1] - start[k];
  for (int i = lane; i < cnt; i += 32) {
    const int beg = start[k] + (int)(len * i / cnt), end = start[k] + (int)(len * (i + 1) / cnt);
    items[n_it[k] + i] = make_int4(k, beg, end, 0);
  }
  // unused tail items: empty
  for (int i = n_it[PB_N_DISTS] + threadIdx.x; i < n_items; i += blockDim.x) items[i] = make_int4(0, 0, 0, 0);
}

__global__ void __launch_bounds__(kDistBlock) dist_scatter_kernel(const int4* __restrict__ out_rec, int64_t n_edges,
                                                                  const int* __restrict__ block_off,
                                                                  const int* __restrict__ dist_ptr,
                                                                  int* __restrict__ dist_perm) {
  __shared__ int warp_cnt[32][PB_N_DISTS + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t p = (int64_t)blockIdx.x * kDistBlock + threadIdx.x;
  const bool ok = p < n_edges;
  const int k = ok ? ((out_rec[p].y >> 8) & 31) : -1;
  for (int i = lane; i < PB_N_DISTS; i += 32) warp_cnt[w][i] = 0;
  __syncwarp();
  const uint32_t same = __match_any_sync(0xffffffffu, k);
  const int rank_in_warp = __popc(same & ((1u << lane) - 1u));
  if (ok && rank_in_warp == 0) warp_cnt[w][k] = __popc(same);
  __syncthreads();
  // exclusive prefix over warps for my distance
  int before = 0;
  if (ok)
    for (int j = 0; j < w; ++j) before += warp_cnt[j][k];
  if (ok) dist_perm[dist_ptr[k] + block_off[(size_t)blockIdx.x * PB_N_DISTS + k] + before + rank_in_warp] = (int)p;
}

struct CsrWs {
  int* seg_cnt;   // [N*R]  counts, then slot cursors
  int* out_cnt;   // [N]
  int* out_eid;   // [E]
  int* block_hist;  // [ceil(E/1024)][32]
  void* scan0;
  void* scan1;
  size_t total;
};

static CsrWs carve(void* base, int64_t n, int64_t e, int r) {
  CsrWs w;
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* q = p ? p + off : nullptr;
    off += align_up(bytes ? bytes : 1, 256);
    return q;
  };
  w.seg_cnt = reinterpret_cast<int*>(take((size_t)n * r * sizeof(int)));
  w.out_cnt = reinterpret_cast<int*>(take((size_t)n * sizeof(int)));
  w.out_eid = reinterpret_cast<int*>(take((size_t)e * sizeof(int)));
  w.block_hist = reinterpret_cast<int*>(take((size_t)((e + kDistBlock - 1) / kDistBlock) * PB_N_DISTS * sizeof(int)));
  w.scan0 = take(scan_workspace_bytes(n * r));
  w.scan1 = take(scan_workspace_bytes(n));
  w.total = off;
  return w;
}

}  // namespace pb

using namespace pb;

extern "C" size_t pb_csr_workspace_bytes(int64_t n_nodes, int64_t n_edges, int32_t n_relations) {
  if (n_nodes < 0 || n_edges < 0 || n_relations <= 0) return 0;
  return carve(nullptr, n_nodes, n_edges, n_relations).total;
}

extern "C" int32_t pb_csr_num_dist_items(void) { return PB_DIST_ITEMS; }

extern "C" int pb_csr_build(const int64_t* edge_index, const uint8_t* edge_type, const uint8_t* edge_dist,
                            int64_t n_nodes, int64_t n_edges, int32_t n_relations, int32_t* in_ptr,
                            int32_t* in_edge, int32_t* in_eid, int32_t* out_ptr, void* out_rec, int32_t* dist_perm,
                            void* dist_items, int32_t* dist_item_ptr, void* workspace, size_t workspace_bytes,
                            pb_stream_t stream) {
  PB_REQUIRE(edge_index && edge_type && edge_dist && in_ptr && in_edge && in_eid && out_ptr && out_rec && workspace &&
                 dist_perm && dist_items && dist_item_ptr,
             "pb_csr_build: null pointer");
  PB_REQUIRE(n_nodes > 0 && n_nodes < ((int64_t)1 << 26), "pb_csr_build: n_nodes=%lld must be in (0, 2^26)",
             (long long)n_nodes);
  PB_REQUIRE(n_edges >= 0 && n_edges < ((int64_t)1 << 31), "pb_csr_build: n_edges out of range");
  PB_REQUIRE(n_relations > 0 && n_relations <= 255 && n_nodes * n_relations < ((int64_t)1 << 31),
             "pb_csr_build: n_relations out of range");
  PB_REQUIRE((reinterpret_cast<uintptr_t>(out_rec) & 15) == 0, "pb_csr_build: out_rec must be 16B aligned");
  CsrWs w = carve(workspace, n_nodes, n_edges, n_relations);
  PB_REQUIRE(workspace_bytes >= w.total, "pb_csr_build: workspace too small (%zu < %zu)", workspace_bytes, w.total);
  cudaStream_t st = as_stream(stream);
  const long long* src = reinterpret_cast<const long long*>(edge_index);
  const long long* dst = src + n_edges;
  const int64_t n_seg = n_nodes * n_relations;
  const int threads = 256;
  const unsigned cap = (unsigned)sm_count() * 16;
  auto grid_for = [&](int64_t n) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, cap)); };

  PB_CUDA(cudaMemsetAsync(w.seg_cnt, 0, (size_t)n_seg * sizeof(int), st));
  PB_CUDA(cudaMemsetAsync(w.out_cnt, 0, (size_t)n_nodes * sizeof(int), st));
  if (n_edges > 0) {
    csr_count_kernel<<<grid_for(n_edges), threads, 0, st>>>(src, dst, edge_type, edge_dist, n_nodes, n_edges,
                                                            n_relations, w.seg_cnt, w.out_cnt);
    PB_LAUNCH_CHECK();
  }
  int rc = exclusive_scan_i32(w.seg_cnt, in_ptr, n_seg, w.scan0, st);
  if (rc) return rc;
  rc = exclusive_scan_i32(w.out_cnt, out_ptr, n_nodes, w.scan1, st);
  if (rc) return rc;
  if (n_edges == 0) {
    PB_CUDA(cudaMemsetAsync(dist_item_ptr, 0, (PB_N_DISTS + 1) * sizeof(int), st));
    PB_CUDA(cudaMemsetAsync(dist_items, 0, (size_t)PB_DIST_ITEMS * sizeof(int4), st));
    return PB_OK;
  }
  PB_CUDA(cudaMemsetAsync(w.seg_cnt, 0, (size_t)n_seg * sizeof(int), st));
  PB_CUDA(cudaMemsetAsync(w.out_cnt, 0, (size_t)n_nodes * sizeof(int), st));
  csr_slot_kernel<<<grid_for(n_edges), threads, 0, st>>>(src, dst, edge_type, edge_dist, n_nodes, n_edges,
                                                         n_relations, in_ptr, out_ptr, w.seg_cnt, w.out_cnt, in_eid,
                                                         w.out_eid);
  PB_LAUNCH_CHECK();
  csr_finish_in_kernel<<<grid_for(n_seg), threads, 0, st>>>(src, edge_dist, n_seg, in_ptr, in_eid, in_edge);
  PB_LAUNCH_CHECK();
  csr_finish_out_kernel<<<grid_for(n_nodes), threads, 0, st>>>(dst, edge_type, edge_dist, n_nodes, n_relations,
                                                               in_ptr, out_ptr, w.out_eid, reinterpret_cast<int4*>(out_rec));
  PB_LAUNCH_CHECK();
  // distance grouping of the out-edge positions. dist_ptr lives in the first 33 ints of scan1's block (free by now).
  const int64_t n_blocks = (n_edges + kDistBlock - 1) / kDistBlock;
  int* dist_ptr = reinterpret_cast<int*>(w.scan1);
  const int4* recs = reinterpret_cast<const int4*>(out_rec);
  dist_hist_kernel<<<(unsigned)n_blocks, kDistBlock, 0, st>>>(recs, n_edges, w.block_hist);
  PB_LAUNCH_CHECK();
  dist_offsets_kernel<<<1, 1024, 0, st>>>(w.block_hist, n_blocks, n_edges, PB_DIST_ITEMS, dist_ptr, dist_item_ptr,
                                          reinterpret_cast<int4*>(dist_items));
  PB_LAUNCH_CHECK();
  dist_scatter_kernel<<<(unsigned)n_blocks, kDistBlock, 0, st>>>(recs, n_edges, w.block_hist, dist_ptr, dist_perm);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
