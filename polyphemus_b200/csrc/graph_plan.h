// Closed-form per-bar graph plan shared by the device kernels (graph_build.cu) and the host-compiled
// logic check in tests/ (compiled with g++; no CUDA needed). Pure integer bit arithmetic on the four
// 32-bit track masks of one bar.
//
// Replaces the Python list building of the reference:
//   node labels   data.py:14-21    (rank of (track, t) in row-major nonzero order)
//   TRACK edges   data.py:24-51    (per track: forward edges between consecutive activations, then inverses)
//   ONSET edges   data.py:54-80    (per timestep: lexicographic track pairs, forward then inverses, dist 0)
//   NEXT edges    data.py:83-121   (consecutive active timesteps, cross-track product, forward only)
//   concat order  data.py:159-167, fake self-edge data.py:173-176
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD inline
#endif

namespace pb {

PB_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
PB_HD int ctz32(uint32_t v) {  // v != 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)v) - 1;
#else
  return __builtin_ctz(v);
#endif
}
PB_HD uint32_t low_mask(int t) { return t >= 32 ? 0xFFFFFFFFu : ((1u << t) - 1u); }
PB_HD uint32_t above(uint32_t v, int t) { return t >= 31 ? 0u : (v >> (t + 1)); }  // bits strictly above t, shifted

struct BarPlan {
  uint32_t b[4];      // track masks
  int n[4];           // nodes per track
  int pre[4];         // label of the first node of each track
  int toff[4];        // offset of each track's TRACK-edge block
  int n_nodes;
  int n_track_edges, n_onset_edges, n_next_edges;
  int n_edges;        // including the fake self-edge for an edgeless bar
};

PB_HD uint32_t column(const uint32_t b[4], int t) {
  return ((b[0] >> t) & 1u) | (((b[1] >> t) & 1u) << 1) | (((b[2] >> t) & 1u) << 2) | (((b[3] >> t) & 1u) << 3);
}

// number of NEXT edges leaving timestep t towards the following active timestep t2
PB_HD int next_pairs(uint32_t col1, uint32_t col2) { return popc32(col1) * popc32(col2) - popc32(col1 & col2); }

PB_HD BarPlan make_bar_plan(const uint32_t bits[4]) {
  BarPlan p;
  int acc = 0, eacc = 0;
  for (int k = 0; k < 4; ++k) {
    p.b[k] = bits[k];
    p.n[k] = popc32(bits[k]);
    p.pre[k] = acc;
    p.toff[k] = eacc;
    acc += p.n[k];
    eacc += p.n[k] > 1 ? 2 * (p.n[k] - 1) : 0;
  }
  p.n_nodes = acc;
  p.n_track_edges = eacc;
  int pairs = 0;
  for (int a = 0; a < 4; ++a)
    for (int c = a + 1; c < 4; ++c) pairs += popc32(bits[a] & bits[c]);
  p.n_onset_edges = 2 * pairs;
  uint32_t u = bits[0] | bits[1] | bits[2] | bits[3];
  int nx = 0;
  while (u) {
    int t1 = ctz32(u);
    u &= u - 1;
    if (!u) break;
    int t2 = ctz32(u);
    nx += next_pairs(column(bits, t1), column(bits, t2));
  }
  p.n_next_edges = nx;
  int e = p.n_track_edges + p.n_onset_edges + p.n_next_edges;
  p.n_edges = e > 0 ? e : 1;
  return p;
}

PB_HD int node_label(const BarPlan& p, int k, int t) { return p.pre[k] + popc32(p.b[k] & low_mask(t)); }

// offset (inside the bar's edge list) of the ONSET block of timestep t
PB_HD int onset_base(const BarPlan& p, int t) {
  int pairs = 0;
  const uint32_t m = low_mask(t);
  for (int a = 0; a < 4; ++a)
    for (int c = a + 1; c < 4; ++c) pairs += popc32(p.b[a] & p.b[c] & m);
  return p.n_track_edges + 2 * pairs;
}

// offset of the NEXT block that starts at active timestep t
PB_HD int next_base(const BarPlan& p, int t) {
  uint32_t u = (p.b[0] | p.b[1] | p.b[2] | p.b[3]) & low_mask(t + 1);  // active timesteps <= t
  int acc = 0;
  while (u) {
    int t1 = ctz32(u);
    u &= u - 1;
    if (!u) break;  // t1 == t
    int t2 = ctz32(u);
    acc += next_pairs(column(p.b, t1), column(p.b, t2));
  }
  return p.n_track_edges + p.n_onset_edges + acc;
}

// Emits every edge whose *source timestep* is t through `emit(pos, u, v, type, dist)` (labels local to
// the bar). One call per timestep covers the whole bar exactly once (plus `emit_fake` below).
template <class Emit>
PB_HD void emit_timestep_edges(const BarPlan& p, int t, Emit emit) {
  const uint32_t col = column(p.b, t);
  if (!col) return;
  // TRACK: forward edge r of track k lives at toff[k] + r, its inverse at toff[k] + (n[k]-1) + r
  for (int k = 0; k < 4; ++k) {
    if (!((col >> k) & 1u)) continue;
    uint32_t rest = above(p.b[k], t);
    if (!rest) continue;
    int dist = 1 + ctz32(rest);
    int r = popc32(p.b[k] & low_mask(t));
    int u = p.pre[k] + r;
    emit(p.toff[k] + r, u, u + 1, k, dist);
    emit(p.toff[k] + (p.n[k] - 1) + r, u + 1, u, k, dist);
  }
  // ONSET: lexicographic pairs (a < c) of active tracks; forward block then inverse block
  int c_t = popc32(col);
  if (c_t >= 2) {
    int base = onset_base(p, t), np = c_t * (c_t - 1) / 2, i = 0;
    for (int a = 0; a < 4; ++a) {
      if (!((col >> a) & 1u)) continue;
      for (int c = a + 1; c < 4; ++c) {
        if (!((col >> c) & 1u)) continue;
        int u = node_label(p, a, t), v = node_label(p, c, t);
        emit(base + i, u, v, 4, 0);
        emit(base + np + i, v, u, 4, 0);
        ++i;
      }
    }
  }
  // NEXT: towards the following active timestep of the bar (any track), same-track pairs excluded
  uint32_t rest = above(p.b[0] | p.b[1] | p.b[2] | p.b[3], t);
  if (rest) {
    int t2 = t + 1 + ctz32(rest);
    uint32_t col2 = column(p.b, t2);
    int pos = next_base(p, t);
    for (int a = 0; a < 4; ++a) {
      if (!((col >> a) & 1u)) continue;
      for (int c = 0; c < 4; ++c) {
        if (!((col2 >> c) & 1u) || c == a) continue;
        emit(pos++, node_label(p, a, t), node_label(p, c, t2), 5, t2 - t);
      }
    }
  }
}

PB_HD bool bar_is_edgeless(const BarPlan& p) {
  return p.n_track_edges + p.n_onset_edges + p.n_next_edges == 0;
}

}  // namespace pb
