/*
 * polyphemus_b200.h — C ABI of libpolyphemus_b200.so (hand-written sm_100a CUDA).
 *
 * Drop-in boundary for ONE hot path of EmanueleCosenza/polyphemus: batched pianoroll-structure -> graph
 * construction (reference data.py:14-204) and the relational graph-convolution stack of its graph VAE
 * (reference model.py:30-135 GCL, model.py:167-208 GCN), forward and backward.
 *
 * The reference is pure Python and has no FFI; each entry point below names the reference lines it
 * replaces. INTEGRATION.md shows the ctypes stub a maintainer would add to the reference.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter is documented as host;
 *   - the caller owns all memory; the library never allocates device memory (workspace sizes come from the
 *     *_workspace_bytes queries) and never synchronises the stream;
 *   - every function enqueues on `stream` (a cudaStream_t passed as void*) and returns 0 on success or a
 *     negative pb_status; pb_last_error() returns a thread-local message for the last failure;
 *   - row-major everywhere; `ld*` are leading dimensions in ELEMENTS;
 *   - integer outputs are bit-exact with the reference; floating-point outputs follow DESIGN.md tolerances;
 *   - there is NO CPU fallback: calls fail with PB_ERR_CUDA when no sm_100 device/context is usable.
 */
#ifndef POLYPHEMUS_B200_H
#define POLYPHEMUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_VERSION 100  /* 0.1.0 */

typedef void* pb_stream_t; /* cudaStream_t */

enum pb_status {
  PB_OK = 0,
  PB_ERR_INVALID = -1,   /* bad argument (shape, alignment, null pointer) */
  PB_ERR_CUDA = -2,      /* CUDA runtime / driver error, message in pb_last_error() */
  PB_ERR_WORKSPACE = -3, /* workspace too small */
  PB_ERR_UNSUPPORTED = -4
};

/* Arithmetic mode of the tensor-core contractions (DESIGN.md "Precision modes").
 *   PB_F32  : operands split into TF32 hi+lo, three tcgen05 kind::tf32 MMAs per k-step, fp32 accumulate
 *             (fp32-grade: the parity mode, rtol 1e-4 / atol 1e-5 against the fp32 reference);
 *   PB_BF16 : operands rounded to bf16, one tcgen05 kind::f16 MMA per k-step, fp32 accumulate. */
enum pb_dtype { PB_F32 = 0, PB_BF16 = 1 };

enum { PB_N_TRACKS = 4, PB_N_TIMESTEPS = 32, PB_N_DISTS = 32, PB_N_RELATIONS = 6, PB_DIST_ITEMS = 1184 };

/* Row groups of the structured node layout: nodes sorted by the relation of their incoming TRACK
 * edges, every group padded with zero rows to a multiple of 128 so that a GEMM tile never mixes two groups.
 * `count[g]` real rows start at padded row `start[g]`. NULL wherever accepted = one group {0, m}. */
typedef struct pb_groups {
  int32_t n_groups; /* <= 4 */
  int32_t reserved;
  int64_t start[4];
  int64_t count[4];
} pb_groups_t;

int pb_version(void);
const char* pb_last_error(void);
/* Host-side query: SM count and compute capability of the current device. */
int pb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * Graph construction  —  replaces data.py:141-204 graph_from_tensor (+ get_*_edges data.py:14-138) and
 * the per-sequence loop + Batch.from_data_list of model.py:596-607 / train.py:152-156.
 * ---------------------------------------------------------------------------------------------- */

/* Pass 1. s_tensor: uint8/bool [n_bars, 4, 32] (all bars of all sequences, sequence-major). Empty bars get
 * the fake activation s[bar,0,0]=1 IN PLACE (data.py:152-153). Outputs:
 *   bar_bits  u32 [n_bars,4]   bit t of word k = activation (track k, timestep t)
 *   node_ptr  i32 [n_bars+1]   exclusive scan of nodes per bar
 *   edge_ptr  i32 [n_bars+1]   exclusive scan of edges per bar (fake self-edge counted, data.py:173-176)
 *   totals    i64 [8]          {N, E, n_drum_nodes, n_bars, nodes per track-relation group g = 0..3}
 *                              (group of a node = the relation of its incoming TRACK edges: its track, or 0 for
 *                              the only node of a one-node bar, whose single in-edge is the fake self-edge)       */
size_t pb_graph_workspace_bytes(int64_t n_bars);
int pb_graph_count(uint8_t* s_tensor, int64_t n_bars, uint32_t* bar_bits, int32_t* node_ptr,
                   int32_t* edge_ptr, int64_t* totals, void* workspace, size_t workspace_bytes,
                   pb_stream_t stream);

/* Pass 2. Emits the reference's arrays in the reference's order (bit-exact):
 *   edge_index    i64 [2,E]  global node ids (PyG collate increment)      data.py:173,193 / model.py:604
 *   edge_type     u8  [E]    0..3 track, 4 onset, 5 next                  constants.py:52-58
 *   edge_dist     u8  [E]    timestep distance 0..31                      data.py:45,74,116
 *   edge_attrs    f32 [E,33] optional (NULL to skip): col0=type, col 1+dist=1   data.py:179-182
 *   node_features f32 [N,4]  one-hot track                                data.py:124-138
 *   is_drum       u8  [N]                                                 data.py:185
 *   bars          i64 [N]    bar index inside its sequence                data.py:202
 *   batch         i64 [N]    sequence index                               PyG add_batch
 *   node_track    u8  [N]    track of the node (internal helper, optional NULL)
 *   node_group    u8  [N]    track-relation group of the node (internal helper, optional NULL)           */
int pb_graph_fill(const uint32_t* bar_bits, const int32_t* node_ptr, const int32_t* edge_ptr,
                  int64_t n_bars, int32_t bars_per_seq, int64_t* edge_index, int64_t n_edges,
                  uint8_t* edge_type, uint8_t* edge_dist, float* edge_attrs, float* node_features,
                  uint8_t* is_drum, int64_t* bars, int64_t* batch, uint8_t* node_track, uint8_t* node_group,
                  pb_stream_t stream);

/* edge_attrs f32 [E,33] from (type, dist)  — data.py:179-182, materialised lazily. */
int pb_edge_attrs_encode(const uint8_t* edge_type, const uint8_t* edge_dist, int64_t n_edges,
                         float* edge_attrs, pb_stream_t stream);
/* Inverse, for foreign graphs handed to GCN/GCL: edge_type f32 [E] (stride in elements) and one-hot
 * edge_attr f32 [E,32] (row stride in elements)  ->  u8 type, u8 dist (argmax).  model.py:193-194     */
int pb_edge_attrs_decode(const float* edge_type, int64_t type_stride, const float* edge_attr,
                         int64_t attr_stride, int64_t n_edges, uint8_t* type_out, uint8_t* dist_out,
                         pb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * CSR plan — replaces the per-relation boolean compaction masked_edge_index/attrs (model.py:30-38,
 * 104-105) and PyG's gather/scatter bookkeeping. Destination-sorted segments keyed by (dst, relation)
 * for the forward mean-aggregation; source-sorted records for the backward scatter. Deterministic:
 * edges inside a segment are ordered by their index in edge_index.
 *   in_ptr  i32 [N*R+1]   in_edge i32 [E] = src | dist<<26     in_eid i32 [E] original edge id
 *   out_ptr i32 [N+1]     out_rec int4 [E] = {dst, rel | dist<<8, eid, |segment(dst,rel)|}
 * plus a grouping of the source-view positions by timestep distance, used by the deterministic reduction of the
 * edge-network gradient: dist_perm i32 [E] (positions grouped by distance, stable), dist_items int4
 * [PB_DIST_ITEMS] = {dist, begin, end, 0} (contiguous, similarly sized slices of dist_perm), dist_item_ptr i32 [33].
 * ---------------------------------------------------------------------------------------------- */
typedef struct pb_csr {
  int64_t n_nodes;
  int64_t n_edges;
  int32_t n_relations;
  int32_t reserved;
  const int32_t* in_ptr;
  const int32_t* in_edge;
  const int32_t* in_eid;
  const int32_t* out_ptr;
  const void* out_rec;
  const int32_t* dist_perm;
  const void* dist_items;
  const int32_t* dist_item_ptr;
  /* optional visiting order of the n_nodes rows (NULL = 0..n-1). The structured layout stores nodes sorted by
   * track relation but visits them bar by bar, so that a bar's rows (spread over the four groups) are gathered
   * while they are still in L2. */
  const int32_t* node_order;
  /* int4 [n_nodes] = {row, first out-edge, out-degree, first in-edge} per visited node, in visiting order
   * (pb_csr_visit_meta, after node_order is set): the fused backward reads it 32 sources per coalesced load instead
   * of chasing node_order -> out_ptr per source. */
  const void* visit_meta;
  /* the fused backward's work list (pb_csr_bwd_stream): int4 [3 n_nodes + n_edges] records in visiting order, per
   * source {x row, root-block row, residual row, out-edge rows...}; visit_edge_ptr i32 [n_nodes + 1] = exclusive sum of
   * the out-degrees in visiting order (record position of source i = 3 i + visit_edge_ptr[i]). */
  const void* bwd_stream;
  const int32_t* visit_edge_ptr;
} pb_csr_t;

size_t pb_csr_workspace_bytes(int64_t n_nodes, int64_t n_edges, int32_t n_relations);
int32_t pb_csr_num_dist_items(void);
int pb_csr_visit_meta(const pb_csr_t* csr, void* visit_meta /* int4 [n_nodes] */, pb_stream_t stream);
int pb_csr_bwd_stream(const pb_csr_t* csr, const int32_t* visit_edge_ptr, void* bwd_stream, pb_stream_t stream);
int pb_csr_build(const int64_t* edge_index, const uint8_t* edge_type, const uint8_t* edge_dist,
                 int64_t n_nodes, int64_t n_edges, int32_t n_relations, int32_t* in_ptr, int32_t* in_edge,
                 int32_t* in_eid, int32_t* out_ptr, void* out_rec, int32_t* dist_perm, void* dist_items,
                 int32_t* dist_item_ptr, void* workspace, size_t workspace_bytes, pb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Edge network table — replaces nn.Linear(32,d) applied to one-hot distances (model.py:127-129, 175):
 * T[k,:] = nn_weight[:,k] + nn_bias.   Backward (fixed-order reduction of the per-item partials of pb_agg_bwd,
 * f32 [PB_DIST_ITEMS, d], items of distance k = [dist_item_ptr[k], dist_item_ptr[k+1])):
 * g_nn_weight[c,k] = dT[k,c], g_nn_bias[c] = sum_k dT[k,c].
 * ---------------------------------------------------------------------------------------------- */
int pb_edge_table_fwd(const float* nn_weight, const float* nn_bias, int32_t d, float* table,
                      pb_stream_t stream);
int pb_edge_table_bwd(const float* dtable_partials, const int32_t* dist_item_ptr, int32_t d, float* g_nn_weight,
                      float* g_nn_bias, pb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Message + mean aggregation — replaces GCL.message (model.py:123-135) and propagate/scatter-mean per
 * relation (model.py:103-111):  H[v,r,:] = mean_{e in seg(v,r)} keep_e * relu(x[src_e] * T[dist_e]).
 * Writes the GEMM operand  A[v, :] = [H[v,0] | ... | H[v,R-1] | x[v]]  (row stride lda):
 *   PB_BF16: A bf16;   PB_F32: A_hi / A_lo fp32 (TF32 split, A_lo may not be NULL).
 * Dropout (model.py:133, p hard-wired 0.1 in training): the keep decision of every (edge, channel) is drawn once
 * per layer call by pb_dropout_bits — SplitMix64 hash of (seed; edge id, channel/4), four 16-bit lanes per
 * 4-channel chunk, keep iff lane >= round(p * 65536) — packed 1 bit per channel, and read by forward, operand
 * recompute and backward alike (keep_bits may be NULL when p_drop == 0). pb_dropout_mask exposes the same
 * decisions as bytes [E, d] for checking.
 * Backward, two kernels, no atomics:
 *   (1) one warp per source node u: gx[u] = gy_res[u] + dA[u, R*d:] + sum_{e: src=u} ds_e * T[dist_e] with
 *       ds_e = dH[dst_e, rel_e]/cnt * keep * 1[x*T>0]; the table-gradient row q_e = ds_e * x[u] of every out-edge
 *       position goes to q_buf [E, d] (bf16 for PB_BF16, f32 for PB_F32);
 *   (2) the rows of q_buf are summed per timestep distance through dist_perm / dist_items (fixed order) into
 *       dtable_partials f32 [PB_DIST_ITEMS, d], which pb_edge_table_bwd finishes.
 * ---------------------------------------------------------------------------------------------- */
/* act_dtype (PB_F32 | PB_BF16) = storage of the activations that travel between the kernels of a stack: node features
 * x / y, the pre-BatchNorm GEMM output `out`, and the gradients gy / gx. PB_F32 is the API dtype (reference tensors are
 * fp32); PB_BF16 is what the throughput mode's GCN stacks keep between layers (the reference holds them in fp16 under
 * autocast, training.py:137). Arithmetic is fp32 in both; row strides are in elements. */
size_t pb_dropout_bits_bytes(int64_t n_edges, int32_t d);
int pb_dropout_bits(int64_t n_edges, int32_t d, float p_drop, uint64_t seed, void* keep_bits, pb_stream_t stream);
int pb_agg_fwd(const pb_csr_t* csr, const void* x, int32_t d, const float* table, void* a_hi, void* a_lo,
               int64_t lda, int32_t dtype, const void* keep_bits, float p_drop, int32_t act_dtype,
               pb_stream_t stream);
int pb_agg_bwd(const pb_csr_t* csr, const void* x, int32_t d, const float* table, const void* d_a,
               int64_t ldda, int32_t dtype, const void* gy_res, void* gx, void* q_buf, float* dtable_partials,
               const void* keep_bits, float p_drop, int32_t act_dtype, pb_stream_t stream);
/* Fused backward (the product path): the same gx and edge-table gradient WITHOUT the q_buf round trip. A CTA owns a
 * contiguous range of sources and its threads own channel columns — of the gathered rows, of gx, and of a [32, d] fp32
 * accumulator in shared memory, so the 32-bin reduction dT[dist_e] += ds_e * x[src_e] needs neither atomics nor
 * barriers and adds in edge order (bit-reproducible). dtable_partials is f32 [pb_agg_bwd_num_partials(n, d), 32, d]
 * (one block per CTA), finished by pb_edge_table_bwd_fused. Needs csr->bwd_stream; d_a must be packed (ldda == (R+1) d). */
int32_t pb_agg_bwd_num_partials(int64_t n_nodes, int32_t d, int32_t dtype);
int pb_agg_bwd_fused(const pb_csr_t* csr, const void* x, int32_t d, const float* table, const void* d_a,
                     int64_t ldda, int32_t dtype, const void* gy_res, void* gx, float* dtable_partials,
                     const void* keep_bits, float p_drop, int32_t act_dtype, pb_stream_t stream);
int pb_edge_table_bwd_fused(const float* dtable_partials, int32_t n_partials, int32_t d, float* g_nn_weight,
                            float* g_nn_bias, pb_stream_t stream);
int pb_dropout_mask(int64_t n_edges, int32_t d, float p_drop, uint64_t seed, uint8_t* keep /*[E,d]*/,
                    pb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Per-relation weight transform — replaces the 7 matmuls of GCL.forward (model.py:112,116,119):
 *   out = A @ Wcat + bias,  Wcat = [weight[0]; ...; weight[R-1]; root]  ((R+1)d x d).
 * tcgen05 (TMEM accumulators, TMA operand loads). Weight operands are prepared once per call by
 * pb_weight_prep (transposed copy for the forward, dtype conversion / TF32 split).
 * ---------------------------------------------------------------------------------------------- */
int pb_weight_prep(const float* weight, const float* root, int32_t n_relations, int32_t d, int32_t dtype,
                   void* wcat_hi, void* wcat_lo, void* wcat_t_hi, void* wcat_t_lo, pb_stream_t stream);
/* out f32 [M, d] = A[M,K] @ Wcat[K,d] + bias (bias may be NULL).  wcat_t_* is [d, K] (K contiguous). */
/* Structured form (groups != NULL): A is [M, 4d] = [H_track | H_onset | H_next | x] with the rows in group order;
 * a row of group g contracts against [weight[g]; weight[4]; weight[5]; root] — the same result with 4/7 of the
 * flops, because a node only ever receives TRACK edges of one relation. k is then 4*d while the weight operands
 * keep their full (R+1)*d extent. */
int pb_rgcn_gemm_fwd(const void* a_hi, const void* a_lo, int64_t lda, const void* wcat_t_hi,
                     const void* wcat_t_lo, const float* bias, void* out, int64_t ldo, int64_t m, int32_t d,
                     int32_t k, const pb_groups_t* groups, int32_t dtype, int32_t act_dtype, pb_stream_t stream);
/* The same forward with the BatchNorm batch statistics as a by-product of the epilogue (model.py:202-203 without a second
 * pass over `out`): bn_partials f32 [pb_rgcn_gemm_fwd_bn_partial_rows(m), 2, d] receives, per (128-row tile, 32-row
 * quadrant), the column sums and sums of squares of the STORED output values over the real rows (the padding rows of a
 * group are excluded). Zero-fill it before the call (quadrants beyond m are not written); finish with pb_bn_finalize. */
int64_t pb_rgcn_gemm_fwd_bn_partial_rows(int64_t m);
int pb_rgcn_gemm_fwd_bn(const void* a_hi, const void* a_lo, int64_t lda, const void* wcat_t_hi,
                        const void* wcat_t_lo, const float* bias, void* out, int64_t ldo, int64_t m, int32_t d,
                        int32_t k, const pb_groups_t* groups, int32_t dtype, int32_t act_dtype, float* bn_partials,
                        pb_stream_t stream);
/* dA [M,K] = g[M,d] @ Wcat^T.  g_* is the GEMM-operand copy of the output gradient (bf16, or f32 hi/lo);
 * dA is bf16 (PB_BF16) or f32 (PB_F32). */
int pb_rgcn_gemm_bwd_data(const void* g_hi, const void* g_lo, int64_t ldg, const void* wcat_hi,
                          const void* wcat_lo, void* d_a, int64_t ldda, int64_t m, int32_t d, int32_t k,
                          const pb_groups_t* groups, int32_t dtype, pb_stream_t stream);
/* Generic D[m,n] = A[m,k] @ B[n,k]^T (+ bias[n]) on the same tcgen05 kernel; out is f32 or bf16. Used for the
 * nn.Linear layers adjacent to the path (chord encoder / decoder, model.py:322,525): forward and input gradient.
 * Their weight gradient is pb_rgcn_gemm_bwd_weight (d = out features, k = in features). */
int pb_gemm_nt(const void* a_hi, const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb,
               const float* bias, void* out, int64_t ldo, int64_t m, int32_t n, int32_t k, int32_t dtype,
               int32_t out_bf16, pb_stream_t stream);
/* dWcat f32 [K,d] = A^T @ g (fixed-order split-K over the node dimension, deterministic). */
size_t pb_rgcn_gemm_bwd_weight_workspace_bytes(int64_t m, int32_t d, int32_t k);            /* any dtype */
size_t pb_rgcn_gemm_bwd_weight_workspace_bytes_for(int64_t m, int32_t d, int32_t k, int32_t dtype);   /* PB_BF16 needs no transposed copies */
/* Structured form (groups != NULL, k == 4d): d_wcat is the full [(R+1)d, d]; the split boundaries follow the row
 * groups, the track block of the partials is reduced per group into weight[g], the other blocks over all rows. */
int pb_rgcn_gemm_bwd_weight(const void* a_hi, const void* a_lo, int64_t lda, const void* g_hi,
                            const void* g_lo, int64_t ldg, float* d_wcat, int64_t m, int32_t d, int32_t k,
                            const pb_groups_t* groups, int32_t dtype, void* workspace, size_t workspace_bytes,
                            pb_stream_t stream);
/* Plain-fp32 CUDA-core contraction D[M,N] = A[M,K] @ B[K,N] (+bias) used by the tests to cross-check the
 * tensor-core kernels on device; not on the product path. */
int pb_gemm_f32_check(const float* a, int64_t lda, const float* b, int64_t ldb, const float* bias, float* d_out,
                      int64_t ldd, int64_t m, int32_t n, int32_t k, int32_t trans_a, int32_t trans_b,
                      pb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm(batch statistics over all node rows) + ReLU + residual — replaces GCN.forward lines
 * model.py:198-206:  y = x + relu(gamma * (out - mean) * rstd + beta).
 * stats: mean/var over the m rows (biased var for normalisation, unbiased for running_var, eps 1e-5,
 * momentum 0.1 — torch.nn.BatchNorm1d). In eval mode pass running stats through pb_bn_prepare_eval.
 * ---------------------------------------------------------------------------------------------- */
size_t pb_bn_workspace_bytes(int64_t m, int32_t d);
/* bn_coef f32 [3,d] = {mean, scale = gamma*rstd, beta}; save_mean_rstd f32 [2,d];
 * running_mean/var updated in place when not NULL. */
int pb_bn_stats(const void* out, int64_t ldo, int64_t m, int32_t d, const pb_groups_t* groups, const float* gamma,
                const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                float* save_mean_rstd, float* bn_coef, void* workspace, size_t workspace_bytes, int32_t act_dtype,
                pb_stream_t stream);
/* mean / var / running statistics / bn_coef from unshifted column partials f32 [n_partials, 2, d] over m_valid rows. */
int pb_bn_finalize(const float* partials, int64_t n_partials, int64_t m_valid, int32_t d, const float* gamma,
                   const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                   float* save_mean_rstd, float* bn_coef, pb_stream_t stream);
int pb_bn_prepare_eval(const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, int32_t d, float* bn_coef,
                       pb_stream_t stream);
/* y = x_res + relu((out-mean)*scale + beta)   (x_res may be NULL: y = relu(...)); apply_relu=0 skips the ReLU */
int pb_bn_relu_res_fwd(const void* out, int64_t ldo, const void* x_res, const float* bn_coef,
                       void* y, int64_t m, int32_t d, const pb_groups_t* groups, int32_t apply_relu,
                       int32_t act_dtype, pb_stream_t stream);
/* Backward of y = x + relu(bn(out)) wrt out (training statistics):
 *   g_out f32 [m,d] (+ GEMM-operand copies g_hi/g_lo in `dtype`), g_gamma, g_beta, g_bias(=colsum g_out).
 *   The residual branch gradient is gy itself (consumed by pb_agg_bwd as gy_res). With row groups the padding rows of
 *   g_hi / g_lo are written as zero (the weight-gradient GEMM contracts over all m rows). */
int pb_bn_relu_res_bwd(const void* gy, const void* out, int64_t ldo, const float* gamma,
                       const float* save_mean_rstd, const float* bn_coef, int64_t m, int32_t d,
                       const pb_groups_t* groups, int32_t dtype, void* g_hi, void* g_lo, int64_t ldg, float* g_gamma,
                       float* g_beta, float* g_bias, void* workspace, size_t workspace_bytes, int32_t act_dtype,
                       pb_stream_t stream);
/* Operand conversion for a plain GCL (no BN): g f32 [m,d] -> g_hi/g_lo in `dtype`, and g_bias = colsum. */
int pb_grad_prep(const float* g, int64_t ldg_in, int64_t m, int32_t d, int32_t dtype, void* g_hi, void* g_lo,
                 int64_t ldg, float* g_bias, void* workspace, size_t workspace_bytes, pb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Token cross entropy — replaces nn.CrossEntropyLoss(ignore_index=PAD) on the pitch / duration logits
 * (training.py:100-101, 320-330) for logits kept per un-embedding head ([rows, classes], row stride ld, classes a
 * multiple of 8 (PB_BF16) / 4 (PB_F32), at most 1024 / 512; padding columns hold -inf).
 *   forward : nll[r] = logsumexp(logits[r]) - logits[r, target[r]],  lse[r] = logsumexp;  rows with
 *             target == ignore_index are not read and give nll = lse = 0.
 *   backward: grad[r, c] = (exp(logits[r,c] - lse[r]) - [c == target[r]]) * row_grad[r]   (0 for ignored rows),
 *             written in the logits' own type.
 * The mean over the kept rows is the caller's (sum(nll) / count), as is picking the drum / non-drum head per node
 * (give each head the target with the other head's rows set to ignore_index).
 * ---------------------------------------------------------------------------------------------- */
int pb_ce_fwd(const void* logits, int64_t ld, int32_t dtype, int64_t rows, int32_t classes, const int32_t* target,
              int32_t ignore_index, float* nll, float* lse, pb_stream_t stream);
int pb_ce_bwd(const void* logits, int64_t ld, int32_t dtype, int64_t rows, int32_t classes, const int32_t* target,
              int32_t ignore_index, const float* lse, const float* row_grad, void* grad, int64_t ldg,
              pb_stream_t stream);
/* The same for n_segments heads that are consecutive column blocks (widths[s] columns each) of one logits matrix,
 * one pass over each row: targets[s] / ignore_index[s] per head (host arrays of n_segments entries; targets[s] are
 * device pointers), nll / lse / row_grad are [n_segments, rows]. The backward writes every column of the row. */
#define PB_CE_MAX_SEGMENTS 4
int pb_ce_rows_fwd(const void* logits, int64_t ld, int32_t dtype, int64_t rows, int32_t n_segments,
                   const int32_t* widths, const int32_t* const* targets, const int32_t* ignore_index, float* nll,
                   float* lse, pb_stream_t stream);
int pb_ce_rows_bwd(const void* logits, int64_t ld, int32_t dtype, int64_t rows, int32_t n_segments,
                   const int32_t* widths, const int32_t* const* targets, const int32_t* ignore_index, const float* lse,
                   const float* row_grad, void* grad, int64_t ldg, pb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Chord embedding — replaces ContentEncoder's token embedding + BatchNorm + chord_encoder Linear + ReLU
 * (model.py:355-388). tokens int16 [n_nodes, tok_stride] hold (pitch id, duration id) pairs, slot t at
 * tok_offset + 2t (the dataset's [N, 16, 2] layout with the SOS slot skipped: tok_stride 32, tok_offset 2).
 * tables [2, n_slots, vocab, d] (f32 for PB_F32, bf16 for PB_BF16) is the folded map
 *   T[set, t, token] = BN(emb)[token] @ W_chord[:, t, half]^T  (pitch tokens first, duration tokens from dur_off),
 * set_id[v] != 0 selects the drum tables (tables[1]).
 *   forward : chord[v] = relu(bias + sum_t T[set_v, t, pitch_vt] + T[set_v, t, dur_off + dur_vt])   f32 [n, d]
 *   bwd_prep: operands of dT = onehot^T @ (g * 1[chord > 0]) for pb_rgcn_gemm_bwd_weight (m = n_nodes,
 *             k = n_slots * vocab_padded, d = 2 d): onehot [n, n_slots * vocab_padded] (bf16 / f32, no low part),
 *             gcat [n, 2 d] with the node's set selecting the column block (bf16, or the TF32 hi/lo pair).
 * ---------------------------------------------------------------------------------------------- */
int pb_chord_embed_fwd(const int16_t* tokens, int64_t tok_stride, int32_t tok_offset, int32_t n_slots,
                       const uint8_t* set_id, const void* tables, int32_t dtype, int32_t vocab, int32_t dur_off,
                       int32_t d, const float* bias, float* chord, int64_t ldc, int64_t n_nodes, pb_stream_t stream);
/* Token histograms per table set (0 = non-drum, 1 = drum nodes) over the n_slots (pitch, duration) pairs from tok_offset:
 * counts i64 [2, n_pitch + n_dur] (pitch bins first). They weight the BatchNorm statistics of the folded embedding
 * tables (BatchNorm over Linear(one_hot) rows, model.py:355-376). Exact integer counts, no host read-back. */
int pb_token_hist(const int16_t* tokens, int64_t tok_stride, int32_t tok_offset, int32_t n_slots, const uint8_t* set_id,
                  int64_t n_nodes, int32_t n_pitch, int32_t n_dur, int64_t* counts, pb_stream_t stream);
int pb_chord_embed_bwd_prep(const int16_t* tokens, int64_t tok_stride, int32_t tok_offset, int32_t n_slots,
                            const uint8_t* set_id, int32_t dur_off, int32_t vocab_padded, int32_t d,
                            const float* chord, int64_t ldc, const float* g, int64_t ldg, int32_t dtype, void* onehot,
                            void* gcat_hi, void* gcat_lo, int64_t n_nodes, pb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Per-bar segment operators. The nodes of bar b are rows bar_ptr[b] .. bar_ptr[b+1] (pb_graph_count's node_ptr; at most
 * 128 per bar); one warp per bar, fixed summation order.
 *   pb_bar_pool_*   — PyG GlobalAttention pooling (model.py:335-340, 409): alpha = softmax over the bar of gate,
 *                     out[b] = sum_v alpha_v h[v];  backward gives g_h and g_gate from g_out.
 *   pb_bar_expand_* — x[v] = z[bar(v)] (model.py:542-546) and its segment-sum gradient g_z[b] = sum_v g_x[v].
 * ---------------------------------------------------------------------------------------------- */
int pb_bar_pool_fwd(const float* h, int64_t ldh, const float* gate, const int32_t* bar_ptr, int64_t n_bars, int32_t d,
                    float* alpha, float* out, pb_stream_t stream);
int pb_bar_pool_bwd(const float* h, int64_t ldh, const float* alpha, const int32_t* bar_ptr, int64_t n_bars, int32_t d,
                    const float* g_out, float* g_h, int64_t ldgh, float* g_gate, pb_stream_t stream);
int pb_bar_expand_fwd(const float* z, const int32_t* bar_ptr, int64_t n_bars, int32_t d, float* x, int64_t ldx,
                      pb_stream_t stream);
int pb_bar_expand_bwd(const float* g_x, int64_t ldg, const int32_t* bar_ptr, int64_t n_bars, int32_t d, float* g_z,
                      pb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Row permutation into / out of the structured layout, fused with the activation-storage conversion (f32 <-> bf16):
 *   scatter: dst[pos[v]] = src[v] for v < n, the padding rows of every group of dst [n_rows, d] written as zero;
 *   gather : dst[v] = src[pos[v]].   pos i64 [n] is injective (node -> padded row). Each is the other's gradient.
 * ---------------------------------------------------------------------------------------------- */
int pb_rows_scatter(const void* src, int32_t src_dtype, const int64_t* pos, int64_t n, int32_t d, void* dst,
                    int32_t dst_dtype, int64_t n_rows, const pb_groups_t* groups, pb_stream_t stream);
int pb_rows_gather(const void* src, int32_t src_dtype, const int64_t* pos, int64_t n, int32_t d, void* dst,
                   int32_t dst_dtype, pb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Data formats either side of the path.
 *   pb_dataset_structure / pb_dataset_tokens — replace PolyphemusDataset.__getitem__ (data.py:218-271) for a batch of
 *     samples in the on-disk layout of preprocess.py:210: s_disk u8/bool [B, 4, T], c_disk int16 [B, 4, T, 16, 2],
 *     T = n_bars * 32. Step 1 reorders the structure to [B, n_bars, 4, 32] (data.py:230-231); pb_graph_count then
 *     applies the fake activation of empty bars and yields bar_bits / node_ptr; step 2 copies the 16 (pitch, duration)
 *     pairs of every active cell into tokens int16 [N, 16, 2] in node order — the silence filter of data.py:264-266.
 *     The one-hot expansion (data.py:233-259) is not materialised: token ids feed pb_chord_embed_fwd directly.
 *   pb_mtp_from_logits — replaces utils.mtp_from_logits (utils.py:59-79): mtp [n_cells, n_tok, d_tok] (f32 or bf16,
 *     cells = flattened [B, n_bars, 4, 32]); an active cell (s_tensor != 0) takes the logits of node node_of_cell[cell]
 *     (= number of active cells before it), a silent one the silence pattern (token 0: pitch_eos, others: pitch_pad).
 * ---------------------------------------------------------------------------------------------- */
int pb_dataset_structure(const uint8_t* s_disk, int64_t n_samples, int32_t n_bars, uint8_t* s_tensor,
                         pb_stream_t stream);
int pb_dataset_tokens(const int16_t* c_disk, const uint32_t* bar_bits, const int32_t* node_ptr, int64_t n_samples,
                      int32_t n_bars, int16_t* tokens, pb_stream_t stream);
int pb_mtp_from_logits(const void* c_logits, int64_t ld_node, int32_t dtype, const uint8_t* s_tensor,
                       const int32_t* node_of_cell, int64_t n_cells, int32_t n_tok, int32_t d_tok, int32_t pitch_eos,
                       int32_t pitch_pad, void* mtp, pb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* POLYPHEMUS_B200_H */
