// Dataset-sample decode and generation post-processing on the device (the data formats either side of the path).
//
//   pb_dataset_structure / pb_dataset_tokens  replace PolyphemusDataset.__getitem__ (reference data.py:218-271) on the
//       on-disk sample layout written by preprocess.py:210 — c_tensor int16 [4, T, 16, 2], s_tensor bool [4, T],
//       T = n_bars * 32 — for a whole batch of samples at once: bars-major reshape (data.py:226-231), fake activation
//       of empty bars (through pb_graph_count, data.py:152-153) and the silence filter (data.py:264-266). The note
//       tokens stay token ids (int16 [N, 16, 2], node order); the N x 16 x 230 one-hot of data.py:233-259 is never
//       materialised because the embedding that consumes it is a table lookup (csrc/chord.cu).
//   pb_mtp_from_logits  replaces utils.mtp_from_logits (reference utils.py:59-79): dense [cells, n_tok, d_tok] pianoroll
//       tensor, active cells take their node's logits, silent cells the silence pattern.
// Both are pure HBM-bound copies (64 B per node in; 13.8 KB per cell out), integer / bit exact.
#include "common.cuh"

namespace pb {

// s_disk u8 [B][4][T] -> s_tensor u8 [B][n_bars][4][32]
__global__ void dataset_structure_kernel(const uint8_t* __restrict__ s_disk, int64_t n_samples, int n_bars,
                                         uint8_t* __restrict__ s_tensor) {
  const int64_t total = n_samples * n_bars * PB_N_TRACKS * PB_N_TIMESTEPS;
  const int t_len = n_bars * PB_N_TIMESTEPS;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % PB_N_TIMESTEPS);
    const int trk = (int)((i / PB_N_TIMESTEPS) % PB_N_TRACKS);
    const int bar = (int)((i / (PB_N_TIMESTEPS * PB_N_TRACKS)) % n_bars);
    const int64_t b = i / ((int64_t)PB_N_TIMESTEPS * PB_N_TRACKS * n_bars);
    s_tensor[i] = s_disk[(b * PB_N_TRACKS + trk) * t_len + bar * PB_N_TIMESTEPS + t] != 0;
  }
}

// One warp per bar (lane = timestep): the active cells of the bar, in (track, timestep) order, are nodes
// node_ptr[bar] + rank; each copies its 16 (pitch, duration) pairs = 64 bytes.
__global__ void __launch_bounds__(256) dataset_tokens_kernel(const int16_t* __restrict__ c_disk,
                                                             const uint32_t* __restrict__ bar_bits,
                                                             const int32_t* __restrict__ node_ptr, int64_t n_bars_total,
                                                             int n_bars, int16_t* __restrict__ tokens) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int t_len = n_bars * PB_N_TIMESTEPS;
  for (int64_t bar = warp; bar < n_bars_total; bar += n_warps) {
    const int64_t b = bar / n_bars;
    const int bar_in = (int)(bar - b * n_bars);
    int node = __ldg(node_ptr + bar);
#pragma unroll
    for (int trk = 0; trk < PB_N_TRACKS; ++trk) {
      const uint32_t bits = __ldg(bar_bits + bar * PB_N_TRACKS + trk);
      if ((bits >> lane) & 1u) {
        const int rank = __popc(bits & ((1u << lane) - 1u));
        const uint4* src = reinterpret_cast<const uint4*>(
            c_disk + (((b * PB_N_TRACKS + trk) * t_len + bar_in * PB_N_TIMESTEPS + lane) * 32));
        uint4* dst = reinterpret_cast<uint4*>(tokens + (int64_t)(node + rank) * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = __ldg(src + q);
      }
      node += __popc(bits);
    }
  }
}

// mtp [cells][n_tok][d_tok]: one CTA per cell; node_of_cell[cell] = exclusive count of active cells before it.
template <typename T>
__global__ void __launch_bounds__(256) mtp_kernel(const T* __restrict__ c_logits, int64_t ld_node, const uint8_t* __restrict__ s,
                                                  const int32_t* __restrict__ node_of_cell, int64_t n_cells, int n_tok,
                                                  int d_tok, int pitch_eos, int pitch_pad, T* __restrict__ mtp) {
  const int per = n_tok * d_tok;
  for (int64_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
    T* out = mtp + cell * per;
    if (s[cell]) {
      const T* in = c_logits + (int64_t)__ldg(node_of_cell + cell) * ld_node;
      for (int i = threadIdx.x; i < per; i += blockDim.x) out[i] = in[i];
    } else {
      for (int i = threadIdx.x; i < per; i += blockDim.x) {
        const int tok = i / d_tok, c = i - tok * d_tok;
        out[i] = (T)((tok == 0 ? c == pitch_eos : c == pitch_pad) ? 1.0f : 0.0f);
      }
    }
  }
}

}  // namespace pb

using namespace pb;

extern "C" int pb_dataset_structure(const uint8_t* s_disk, int64_t n_samples, int32_t n_bars, uint8_t* s_tensor,
                                    pb_stream_t stream) {
  PB_REQUIRE(s_disk && s_tensor && n_samples > 0 && n_bars > 0, "pb_dataset_structure: bad arguments");
  const int64_t total = n_samples * n_bars * PB_N_TRACKS * PB_N_TIMESTEPS;
  const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 8);
  dataset_structure_kernel<<<grid, 256, 0, as_stream(stream)>>>(s_disk, n_samples, n_bars, s_tensor);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_dataset_tokens(const int16_t* c_disk, const uint32_t* bar_bits, const int32_t* node_ptr,
                                 int64_t n_samples, int32_t n_bars, int16_t* tokens, pb_stream_t stream) {
  PB_REQUIRE(c_disk && bar_bits && node_ptr && tokens && n_samples > 0 && n_bars > 0, "pb_dataset_tokens: bad arguments");
  PB_REQUIRE((reinterpret_cast<uintptr_t>(c_disk) & 15) == 0 && (reinterpret_cast<uintptr_t>(tokens) & 15) == 0,
             "pb_dataset_tokens: c_disk / tokens must be 16-byte aligned");
  const int64_t bars = n_samples * n_bars;
  const unsigned grid = (unsigned)std::min<int64_t>((bars * 32 + 255) / 256, (int64_t)sm_count() * 16);
  dataset_tokens_kernel<<<grid, 256, 0, as_stream(stream)>>>(c_disk, bar_bits, node_ptr, bars, n_bars, tokens);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_mtp_from_logits(const void* c_logits, int64_t ld_node, int32_t dtype, const uint8_t* s_tensor,
                                  const int32_t* node_of_cell, int64_t n_cells, int32_t n_tok, int32_t d_tok,
                                  int32_t pitch_eos, int32_t pitch_pad, void* mtp, pb_stream_t stream) {
  PB_REQUIRE(c_logits && s_tensor && node_of_cell && mtp && n_cells > 0 && n_tok > 0 && d_tok > 0,
             "pb_mtp_from_logits: bad arguments");
  PB_REQUIRE(ld_node >= (int64_t)n_tok * d_tok, "pb_mtp_from_logits: ld_node smaller than a node's logits");
  PB_REQUIRE(pitch_eos >= 0 && pitch_eos < d_tok && pitch_pad >= 0 && pitch_pad < d_tok, "pb_mtp_from_logits: bad token ids");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_BF16, "pb_mtp_from_logits: bad dtype");
  const unsigned grid = (unsigned)std::min<int64_t>(n_cells, (int64_t)sm_count() * 16);
  if (dtype == PB_F32)
    mtp_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float*>(c_logits), ld_node, s_tensor,
                                                           node_of_cell, n_cells, n_tok, d_tok, pitch_eos, pitch_pad,
                                                           reinterpret_cast<float*>(mtp));
  else
    mtp_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(c_logits), ld_node, s_tensor, node_of_cell, n_cells, n_tok, d_tok, pitch_eos,
        pitch_pad, reinterpret_cast<__nv_bfloat16*>(mtp));
  PB_LAUNCH_CHECK();
  return PB_OK;
}

// ---------------------------------------------------------------------------------------------- token histograms
// Token counts per table set for the BatchNorm statistics of the folded chord embedding (vae.ContentEncoder._bn_table:
// BatchNorm over Linear(one_hot) rows == histogram-weighted statistics of the table rows, model.py:355-376).
// counts i64 [2][n_pitch + n_dur]: set 0 = non-drum nodes, set 1 = drum nodes; pitch bins first, duration bins after.
// Shared-memory histogram per CTA, integer atomics (exact, order-independent), no host read-back.
namespace pb {
__global__ void __launch_bounds__(256) token_hist_kernel(const int16_t* __restrict__ tokens, int64_t tok_stride,
                                                         int tok_offset, int n_slots, const uint8_t* __restrict__ set_id,
                                                         int64_t n_nodes, int n_pitch, int n_dur,
                                                         unsigned long long* __restrict__ counts) {
  extern __shared__ unsigned int hist[];                 // [2][n_pitch + n_dur]
  const int bins = n_pitch + n_dur;
  for (int i = threadIdx.x; i < 2 * bins; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  const int64_t total = n_nodes * n_slots;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / n_slots;
    const int t = (int)(i - v * n_slots);
    const int16_t* tk = tokens + v * tok_stride + tok_offset + 2 * t;
    const int p = tk[0], du = tk[1];
    unsigned int* h = hist + (set_id[v] ? bins : 0);
    if (p >= 0 && p < n_pitch) atomicAdd(h + p, 1u);
    if (du >= 0 && du < n_dur) atomicAdd(h + n_pitch + du, 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * bins; i += blockDim.x)
    if (hist[i]) atomicAdd(counts + i, (unsigned long long)hist[i]);
}
}  // namespace pb

extern "C" int pb_token_hist(const int16_t* tokens, int64_t tok_stride, int32_t tok_offset, int32_t n_slots,
                             const uint8_t* set_id, int64_t n_nodes, int32_t n_pitch, int32_t n_dur, int64_t* counts,
                             pb_stream_t stream) {
  PB_REQUIRE(tokens && set_id && counts && n_nodes >= 0 && n_slots > 0 && n_pitch > 0 && n_dur > 0, "pb_token_hist: bad arguments");
  const int bins = n_pitch + n_dur;
  cudaStream_t st = as_stream(stream);
  PB_CUDA(cudaMemsetAsync(counts, 0, (size_t)2 * bins * sizeof(int64_t), st));
  if (n_nodes == 0) return PB_OK;
  const int64_t total = n_nodes * n_slots;
  const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 4);
  pb::token_hist_kernel<<<grid, 256, (size_t)2 * bins * sizeof(unsigned int), st>>>(
      tokens, tok_stride, tok_offset, n_slots, set_id, n_nodes, n_pitch, n_dur, reinterpret_cast<unsigned long long*>(counts));
  PB_LAUNCH_CHECK();
  return PB_OK;
}
