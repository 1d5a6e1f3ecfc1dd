// Host-compiled check of polyphemus_b200/csrc/graph_plan.h (the integer logic the CUDA graph builder runs
// per (bar, timestep)). TEST INFRASTRUCTURE: lets the CPU-only test tier compare the closed-form edge
// placement with the oracle without a GPU. Never linked into the product library.
#include <stdint.h>
#include "../../polyphemus_b200/csrc/graph_plan.h"

extern "C" int pbh_bar_edges(const uint32_t* bits, int64_t* out /*[n_edges,4]*/, int32_t* counts /*[5]*/) {
  pb::BarPlan p = pb::make_bar_plan(bits);
  counts[0] = p.n_nodes;
  counts[1] = p.n_track_edges;
  counts[2] = p.n_onset_edges;
  counts[3] = p.n_next_edges;
  counts[4] = p.n_edges;
  for (int i = 0; i < p.n_edges * 4; ++i) out[i] = -1;
  for (int t = 0; t < 32; ++t)
    pb::emit_timestep_edges(p, t, [&](int pos, int u, int v, int type, int dist) {
      out[pos * 4 + 0] = u; out[pos * 4 + 1] = v; out[pos * 4 + 2] = type; out[pos * 4 + 3] = dist;
    });
  if (pb::bar_is_edgeless(p)) { out[0] = 0; out[1] = 0; out[2] = 0; out[3] = 0; }
  return p.n_edges;
}
