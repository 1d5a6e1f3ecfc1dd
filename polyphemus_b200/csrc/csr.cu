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

struct CsrWs {
  int* seg_cnt;   // [N*R]  counts, then slot cursors
  int* out_cnt;   // [N]
  int* out_eid;   // [E]
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

extern "C" int pb_csr_build(const int64_t* edge_index, const uint8_t* edge_type, const uint8_t* edge_dist,
                            int64_t n_nodes, int64_t n_edges, int32_t n_relations, int32_t* in_ptr,
                            int32_t* in_edge, int32_t* in_eid, int32_t* out_ptr, void* out_rec, void* workspace,
                            size_t workspace_bytes, pb_stream_t stream) {
  PB_REQUIRE(edge_index && edge_type && edge_dist && in_ptr && in_edge && in_eid && out_ptr && out_rec && workspace,
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
  if (n_edges == 0) return PB_OK;
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
  return PB_OK;
}
