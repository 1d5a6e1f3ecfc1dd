// On-device batched graph construction: replaces reference data.py:14-204 (graph_from_tensor and the
// get_{track,onset,next}_edges list builders) plus the per-sequence loop / Batch.from_data_list of
// model.py:596-607 and train.py:152-156. One warp per bar, one lane per timestep; every edge position is
// closed-form (graph_plan.h), so there are no atomics, no sort and no host round trips except the single
// read-back of {N, E} the caller needs to size its outputs.
#include "common.cuh"
#include "graph_plan.h"
#include "scan.cuh"

namespace pb {

constexpr int kBarsPerBlock = 8;  // 8 warps

// ---- pass 1: pack bits, fake activation (data.py:152-153), per-bar node / edge counts
__global__ void __launch_bounds__(kBarsPerBlock * 32) graph_count_kernel(uint8_t* __restrict__ s_tensor,
                                                                        int64_t n_bars,
                                                                        uint32_t* __restrict__ bar_bits,
                                                                        int* __restrict__ node_cnt,
                                                                        int* __restrict__ edge_cnt,
                                                                        unsigned long long* __restrict__ totals) {
  const int lane = threadIdx.x & 31;
  const int64_t bar = (int64_t)blockIdx.x * kBarsPerBlock + (threadIdx.x >> 5);
  if (bar >= n_bars) return;
  uint8_t* s = s_tensor + bar * 128;
  uint32_t bits[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) bits[k] = __ballot_sync(0xffffffffu, s[k * 32 + lane] != 0);
  if ((bits[0] | bits[1] | bits[2] | bits[3]) == 0u) {
    bits[0] = 1u;
    if (lane == 0) s[0] = 1;
  }
  if (lane == 0) {
    BarPlan p = make_bar_plan(bits);
#pragma unroll
    for (int k = 0; k < 4; ++k) bar_bits[bar * 4 + k] = bits[k];
    node_cnt[bar] = p.n_nodes;
    edge_cnt[bar] = p.n_edges;
    atomicAdd(totals + 2, (unsigned long long)p.n[0]);  // drum nodes (integer add: deterministic)
    // nodes per track-relation group (see node_group in pb_graph_fill): the only node of a one-node bar receives
    // the fake self-edge of type 0 (data.py:173-176), every other node receives track edges of its own track
    if (p.n_nodes == 1) {
      atomicAdd(totals + 4, 1ull);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (p.n[k]) atomicAdd(totals + 4 + k, (unsigned long long)p.n[k]);
    }
  }
}

__global__ void graph_totals_kernel(const int* __restrict__ node_ptr, const int* __restrict__ edge_ptr,
                                    int64_t n_bars, long long* __restrict__ totals) {
  totals[0] = node_ptr[n_bars];
  totals[1] = edge_ptr[n_bars];
  totals[3] = n_bars;
}

// ---- pass 2: fill
__global__ void __launch_bounds__(kBarsPerBlock * 32)
    graph_fill_kernel(const uint32_t* __restrict__ bar_bits, const int* __restrict__ node_ptr,
                      const int* __restrict__ edge_ptr, int64_t n_bars, int bars_per_seq,
                      long long* __restrict__ edge_src, long long* __restrict__ edge_dst,
                      uint8_t* __restrict__ edge_type, uint8_t* __restrict__ edge_dist,
                      float* __restrict__ edge_attrs, float* __restrict__ node_features,
                      uint8_t* __restrict__ is_drum, long long* __restrict__ bars, long long* __restrict__ batch,
                      uint8_t* __restrict__ node_track, uint8_t* __restrict__ node_group) {
  const int lane = threadIdx.x & 31;
  const int64_t bar = (int64_t)blockIdx.x * kBarsPerBlock + (threadIdx.x >> 5);
  if (bar >= n_bars) return;
  uint32_t bits[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) bits[k] = bar_bits[bar * 4 + k];
  const BarPlan p = make_bar_plan(bits);
  const long long nbase = node_ptr[bar];
  const long long ebase = edge_ptr[bar];

  auto emit = [&](int pos, int u, int v, int type, int dist) {
    const long long e = ebase + pos;
    edge_src[e] = nbase + u;
    edge_dst[e] = nbase + v;
    edge_type[e] = (uint8_t)type;
    edge_dist[e] = (uint8_t)dist;
  };
  emit_timestep_edges(p, lane, emit);
  if (lane == 0 && bar_is_edgeless(p)) emit(0, 0, 0, 0, 0);  // data.py:173-176

  // node attributes: lane t owns the (<= 4) nodes of timestep t
  const uint32_t col = column(p.b, lane);
  const long long seq = bar / bars_per_seq, bar_in_seq = bar % bars_per_seq;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!((col >> k) & 1u)) continue;
    const long long v = nbase + node_label(p, k, lane);
    reinterpret_cast<float4*>(node_features)[v] = make_float4(k == 0, k == 1, k == 2, k == 3);
    is_drum[v] = (k == 0);
    bars[v] = bar_in_seq;
    batch[v] = seq;
    if (node_track) node_track[v] = (uint8_t)k;
    if (node_group) node_group[v] = p.n_nodes == 1 ? (uint8_t)0 : (uint8_t)k;
  }

  // optional dense edge_attrs rows (data.py:179-182): the warp writes each 33-float row cooperatively
  if (edge_attrs) {
    __syncwarp();  // the warp's own writes to edge_type/edge_dist above must be visible below
    for (int e = 0; e < p.n_edges; ++e) {
      const long long ge = ebase + e;
      const int type = edge_type[ge], dist = edge_dist[ge];
      float* row = edge_attrs + ge * 33;
      row[lane + 1] = (lane == dist) ? 1.f : 0.f;
      if (lane == 0) row[0] = (float)type;
    }
  }
}

__global__ void edge_attrs_encode_kernel(const uint8_t* __restrict__ type, const uint8_t* __restrict__ dist,
                                         int64_t n_edges, float* __restrict__ attrs) {
  const int64_t total = n_edges * 33;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / 33;
    const int c = (int)(i - e * 33);
    attrs[i] = c == 0 ? (float)type[e] : (c - 1 == dist[e] ? 1.f : 0.f);
  }
}

// warp per edge: argmax over the 32 one-hot columns (first maximum, as torch.argmax), type = (u8) float
__global__ void edge_attrs_decode_kernel(const float* __restrict__ type, int64_t type_stride,
                                         const float* __restrict__ attr, int64_t attr_stride, int64_t n_edges,
                                         uint8_t* __restrict__ type_out, uint8_t* __restrict__ dist_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < n_edges; e += warps) {
    float v = attr[e * attr_stride + lane];
    int idx = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, v, o);
      int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (lane == 0) {
      dist_out[e] = (uint8_t)idx;
      type_out[e] = (uint8_t)(int)type[e * type_stride];
    }
  }
}

}  // namespace pb

using namespace pb;

extern "C" size_t pb_graph_workspace_bytes(int64_t n_bars) {
  if (n_bars < 0) n_bars = 0;
  return 2 * align_up((size_t)n_bars * sizeof(int), 256) + 2 * scan_workspace_bytes(n_bars) + 256;
}

extern "C" int pb_graph_count(uint8_t* s_tensor, int64_t n_bars, uint32_t* bar_bits, int32_t* node_ptr,
                              int32_t* edge_ptr, int64_t* totals, void* workspace, size_t workspace_bytes,
                              pb_stream_t stream) {
  PB_REQUIRE(s_tensor && bar_bits && node_ptr && edge_ptr && totals && workspace, "pb_graph_count: null pointer");
  PB_REQUIRE(n_bars > 0 && n_bars < (int64_t)1 << 24, "pb_graph_count: n_bars=%lld out of range", (long long)n_bars);
  PB_REQUIRE(workspace_bytes >= pb_graph_workspace_bytes(n_bars), "pb_graph_count: workspace too small");
  cudaStream_t st = as_stream(stream);
  char* ws = reinterpret_cast<char*>(workspace);
  int* node_cnt = reinterpret_cast<int*>(ws);
  ws += align_up((size_t)n_bars * sizeof(int), 256);
  int* edge_cnt = reinterpret_cast<int*>(ws);
  ws += align_up((size_t)n_bars * sizeof(int), 256);
  void* scan_ws0 = ws;
  ws += scan_workspace_bytes(n_bars);
  void* scan_ws1 = ws;
  PB_CUDA(cudaMemsetAsync(totals, 0, 8 * sizeof(int64_t), st));
  const unsigned grid = (unsigned)((n_bars + kBarsPerBlock - 1) / kBarsPerBlock);
  graph_count_kernel<<<grid, kBarsPerBlock * 32, 0, st>>>(s_tensor, n_bars, bar_bits, node_cnt, edge_cnt,
                                                          reinterpret_cast<unsigned long long*>(totals));
  PB_LAUNCH_CHECK();
  int rc = exclusive_scan_i32(node_cnt, node_ptr, n_bars, scan_ws0, st);
  if (rc) return rc;
  rc = exclusive_scan_i32(edge_cnt, edge_ptr, n_bars, scan_ws1, st);
  if (rc) return rc;
  graph_totals_kernel<<<1, 1, 0, st>>>(node_ptr, edge_ptr, n_bars, reinterpret_cast<long long*>(totals));
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_graph_fill(const uint32_t* bar_bits, const int32_t* node_ptr, const int32_t* edge_ptr,
                             int64_t n_bars, int32_t bars_per_seq, int64_t* edge_index, int64_t n_edges,
                             uint8_t* edge_type, uint8_t* edge_dist, float* edge_attrs, float* node_features,
                             uint8_t* is_drum, int64_t* bars, int64_t* batch, uint8_t* node_track,
                             uint8_t* node_group, pb_stream_t stream) {
  PB_REQUIRE(bar_bits && node_ptr && edge_ptr && edge_index && edge_type && edge_dist && node_features &&
                 is_drum && bars && batch,
             "pb_graph_fill: null pointer");
  PB_REQUIRE(n_bars > 0 && bars_per_seq > 0 && n_edges > 0, "pb_graph_fill: bad sizes");
  PB_REQUIRE((reinterpret_cast<uintptr_t>(node_features) & 15) == 0, "pb_graph_fill: node_features must be 16B aligned");
  const unsigned grid = (unsigned)((n_bars + kBarsPerBlock - 1) / kBarsPerBlock);
  graph_fill_kernel<<<grid, kBarsPerBlock * 32, 0, as_stream(stream)>>>(
      bar_bits, node_ptr, edge_ptr, n_bars, bars_per_seq, reinterpret_cast<long long*>(edge_index),
      reinterpret_cast<long long*>(edge_index) + n_edges, edge_type, edge_dist, edge_attrs, node_features,
      is_drum, reinterpret_cast<long long*>(bars), reinterpret_cast<long long*>(batch), node_track, node_group);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_edge_attrs_encode(const uint8_t* edge_type, const uint8_t* edge_dist, int64_t n_edges,
                                    float* edge_attrs, pb_stream_t stream) {
  PB_REQUIRE(edge_type && edge_dist && edge_attrs && n_edges >= 0, "pb_edge_attrs_encode: bad arguments");
  if (n_edges == 0) return PB_OK;
  int64_t total = n_edges * 33;
  unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
  edge_attrs_encode_kernel<<<grid, 256, 0, as_stream(stream)>>>(edge_type, edge_dist, n_edges, edge_attrs);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_edge_attrs_decode(const float* edge_type, int64_t type_stride, const float* edge_attr,
                                    int64_t attr_stride, int64_t n_edges, uint8_t* type_out, uint8_t* dist_out,
                                    pb_stream_t stream) {
  PB_REQUIRE(edge_type && edge_attr && type_out && dist_out && n_edges >= 0, "pb_edge_attrs_decode: bad arguments");
  if (n_edges == 0) return PB_OK;
  unsigned grid = (unsigned)std::min<int64_t>((n_edges + 7) / 8, (int64_t)sm_count() * 16);
  edge_attrs_decode_kernel<<<grid, 256, 0, as_stream(stream)>>>(edge_type, type_stride, edge_attr, attr_stride,
                                                                n_edges, type_out, dist_out);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
