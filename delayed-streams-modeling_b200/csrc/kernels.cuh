// Launch wrappers for the hand-written kernels (definitions in mimi_kernels.cu / lm_kernels.cu).
#pragma once

#include "common.cuh"

namespace dsm {

// ---- ring KV bookkeeping (kv_cache.rs:119-237 without the host round trip) -------------------
// Per stream: reads mask[b]; publishes the pre-advance write index / position and the rope base
// position for this step; advances (index, position) by t when active.
struct RingState {
  uint32_t* positions;  // [B] absolute position (can exceed context)
  uint32_t* indices;    // [B] next write slot
  uint32_t* cur_index;  // [B] slot the first new token is written to this step
  uint32_t* cur_pos;    // [B] position of the first new token this step
  uint32_t* rope_pos;   // [B] positions[b] as the reference reads it AFTER the advance
};
void launch_ring_advance(const RingState& rs, const uint8_t* mask, int B, int t, int context, cudaStream_t st);

// The visibility rules every attention kernel (and the known-answer debug entry dsm_debug_ring_trace) shares.  The
// reference materialises an additive mask [B, 1, t, ctx]; here it is recomputed from (write index, position).
#ifdef __CUDACC__
// slot the s-th new token of this step is written to (kv_cache.rs:216-223)
__device__ __forceinline__ int ring_write_slot(int i0, int s, int context) {
  const int w = i0 + s;
  return w >= context ? w - context : w;
}
// t == 1 fast path (kv_cache.rs:130-170: slot j is masked iff j > position): visible slots are [0, n)
__device__ __forceinline__ int ring_visible_count(uint32_t p0, int context) {
  return p0 + 1 < (uint32_t)context ? (int)(p0 + 1) : context;
}
// any t (kv_cache.rs:192-229): slot j holds the position the ring layout implies, overridden for the slots written
// this step; query s (position p0 + s) sees it iff that position <= its own.  An inactive stream's mask is all zero.
__device__ __forceinline__ bool ring_slot_visible(int j, int i0, long long p0, int t, int context, bool active, int s) {
  if (!active) return true;
  long long pj;
  if (p0 < context) pj = j < p0 ? j : (1LL << 40);
  else pj = j < i0 ? j + (p0 - i0) : j + (p0 - i0) - context;
  for (int w = 0; w < t; ++w)
    if (j == ring_write_slot(i0, w, context)) pj = p0 + w;
  return pj <= p0 + s;
}
#endif
// Debug: indices [B, t] and the additive mask [B, t, context] (0 / -inf) of the step ring_advance just published.
void launch_ring_debug(const RingState& rs, const uint8_t* mask, int B, int t, int context, uint32_t* indices, float* amask,
                       cudaStream_t st);

// ---- SEANet ------------------------------------------------------------------------------------
// First encoder conv (C_in = 1): x(t) = t < S ? state[b, t] : pcm[b, t - S]; writes raw and ELU copies.
void launch_init_conv(const float* pcm, const float* state, int S, const float* w /*[C,k]*/, const float* bias,
                      int B, int T, int C, int k, float* out_raw, RowMap raw_map, float* out_act, RowMap act_map,
                      cudaStream_t st);
// SEANet encoder stage 1 in one kernel (seanet_fused.cu): init conv -> resnet block -> ELU.  x1 is the k3 conv's input
// buffer [B][2 + T][64] (head rows read as left context, this frame's last two ELU'd rows written behind them);
// out_act / out_map address the ELU'd block output rows (b, t).
bool seanet_stage1_supported(int c0, int k0, int c1_hidden, int k1, int k2, int T);
void launch_seanet_stage1(const float* pcm, const float* state0, const float* w0b /*[64][8]: 7 taps + bias*/, const float* w1t,
                          const float* b1, const float* w2t, const float* b2, float* x1, long long x1_stride_b, float* out_act,
                          RowMap out_map, int B, int T, cudaStream_t st);
// Last decoder conv (C_out = 1): pcm[b, t] = bias + sum_{j,ci} w[j*C+ci] * x[b, t+j, ci] (x already ELU'd).
void launch_final_conv(const float* x, RowMap x_map /*row = (b, t)*/, const float* w /*[k*C]*/, const float* bias,
                       int B, int T, int C, int k, float* pcm, cudaStream_t st);

// One entry per state-carrying buffer: for active streams copy `count` floats from src to dst.
struct CommitEntry {
  float* dst;
  const float* src;
  long long dst_stride_b, src_stride_b;
  int count;
  int zero_if_inactive_first;  // replicate-padded layer: an inactive stream's state is zeroed on the first step
};
void launch_commit_state(const CommitEntry* entries_dev, int n_entries, int max_count, const uint8_t* mask,
                         int* first_step_flag, int B, cudaStream_t st);
// Replicate left pad on a module's first step (conv.rs:318-327 with PadMode::Replicate): rows [0,S) <- row S.
void launch_replicate_pad(float* x, long long stride_b, int S, int C, const int* first_step_flag, int B,
                          cudaStream_t st);
// Zero every entry's state for streams [b0, b0 + nb) (per-slot reset, conv.rs:274-281; reset_state).
void launch_zero_rows(const CommitEntry* entries_dev, int n_entries, int b0, int nb, cudaStream_t st);

// ---- codec transformer (f32) ---------------------------------------------------------------------
void launch_layernorm(const float* x, RowMap x_map, const float* w, const float* b, float eps, int M, int D,
                      float* out /*dense [M,D]*/, cudaStream_t st, __nv_bfloat16* planes = nullptr,
                      long long plane_stride = 0);   // planes: also the 3-way bf16 split of `out` (see split3_bf16)
// qkv [M, 3*H*D] (M = B*t) -> q rotated in place; rotated k and v scattered into the rings [B,H,ctx,D].
void launch_rope_scatter_f32(float* qkv, const RingState& rs, const uint8_t* mask, int B, int t, int H, int D,
                             int context, const float* inv_freq /*[D/2]*/, float* k_ring, float* v_ring, cudaStream_t st);
// out[m, h*D + d] = softmax(q . K^T * D^-0.5 + mask) . V over the ring, visibility by stored position.
void launch_codec_attention(const float* qkv, const RingState& rs, const uint8_t* mask, int B, int t, int H, int D,
                            int context, const float* k_ring, const float* v_ring, float* out, cudaStream_t st,
                            __nv_bfloat16* planes = nullptr, long long plane_stride = 0);

// ---- RVQ -----------------------------------------------------------------------------------------
// scores [B, bins] = c2 - x.E^T (from the GEMM); idx = first argmin; residual[b,:] -= E[idx,:]; codes[b*n_q+q] = idx.
void launch_rvq_select(const float* scores, int B, int bins, const float* embedding /*[bins,qdim]*/, int qdim,
                       float* residual, long long res_stride, uint32_t* codes, int n_q, int q, cudaStream_t st,
                       __nv_bfloat16* planes = nullptr, long long plane_stride = 0);
// The whole split-RVQ encode as one cluster kernel (rvq_fused.cu): embT [n_q][qdim][bins], c2 [n_q][bins],
// emb [n_q][bins][qdim], resid [B][2*qdim] (first | rest group projections) -> codes [B][n_q].
bool rvq_fused_supported(int bins, int qdim);
void launch_rvq_fused(const float* embT, const float* c2, const float* emb, const float* resid, int B, int n_q, int bins,
                      int qdim, uint32_t* codes, cudaStream_t st);
// decode: out[b, :] = sum_q E_q[codes[b,q], :] over q in [q0, q1)  (quantization.rs:231-248)
// Grouped RVQ encode (rvq_gram.cu): Gram table E_p E_q^T [bins][bins] (built once), and the argmin chain of one group
// of g <= DSM_RVQ_MAX_GROUP codebooks over scores [B][g][bins] of the group's entry residual.
constexpr int DSM_RVQ_MAX_GROUP = 8;
void launch_rvq_gram_table(const float* Ep, const float* Eq, int bins, int qd, float* G, cudaStream_t st);
bool rvq_chain_supported(int bins, int g);
void launch_rvq_chain(const float* scores, int B, int g, int bins, const float* gram, const float* emb /*[g][bins][qd]*/, int qd,
                      float* residual, long long res_stride, uint32_t* codes, int n_q, int q0, __nv_bfloat16* planes,
                      long long plane_stride, cudaStream_t st);
void launch_rvq_gather_sum(const uint32_t* codes, int n_q, int q0, int q1, const float* embeddings /*[n_q][bins][qdim]*/,
                           int bins, int qdim, int B, float* out, cudaStream_t st);
// depthwise conv-transpose k=2s (ConvTrUpsample1d, conv.rs:558-606), one input frame -> s frames,
// carried state = previous input row.  x [B, C]; state [B, C]; out rows (b, r) r < s.
void launch_upsample_dw(const float* x, const float* state, const float* w /*[C, 2s]*/, int B, int C, int s,
                        float* out, RowMap out_map, cudaStream_t st);

// ---- LM (bf16 operands, f32 accumulate) -----------------------------------------------------------
void launch_embed_sum(const uint32_t* text_ids, const uint32_t* audio_ids, int n_cb, const __nv_bfloat16* text_emb,
                      int text_vocab, const __nv_bfloat16* audio_emb /*[n_cb][audio_vocab][D]*/, int audio_vocab,
                      const float* cond, int B, int D, float* x, cudaStream_t st);
// Split-K partial sums of a projection ([count][rows][cols] f32 planes `stride` elements apart, gemm_tc.cuh) that the
// consuming kernel adds in split order; count == 0: the projection wrote its result directly.
struct PartialSums {
  const float* planes = nullptr;
  int count = 0;
  long long stride = 0;
};
// x <- x + sum of the partial planes (if any), then RMSNorm of x -> out_bf16 / out_f32
void launch_rmsnorm(float* x, const PartialSums& ps, const float* alpha, float eps, int B, int D, __nv_bfloat16* out_bf16,
                    float* out_f32, cudaStream_t st);
// qkv f32 [B, 3*H*D] (or its partial planes) -> q bf16 [B, H*D] (rotated); rotated k, v -> bf16 rings [B,H,ctx,D] at
// cur_index[b].
void launch_rope_scatter_bf16(const float* qkv, const PartialSums& ps, const RingState& rs, int B, int H, int D, int context,
                              const float* inv_freq /*[D/2]*/,
                              __nv_bfloat16* q_out, __nv_bfloat16* k_ring, __nv_bfloat16* v_ring, cudaStream_t st);
void launch_lm_attention(const __nv_bfloat16* q, const RingState& rs, const uint8_t* mask, int B, int H, int D,
                         int context, const __nv_bfloat16* k_ring, const __nv_bfloat16* v_ring,
                         __nv_bfloat16* out, cudaStream_t st);
// 8-bit ring of one layer (see lm_kernels.cu): K int8 [B,H,ctx,D], V int8 [B,H,ctxp/4,D,4], scales f32 [B,H,ctxp].
// DSM_KV_INT8K keeps V in bf16 [B,H,ctx,D] (v_bf16 set; v and v_scale unused).
struct KvRingQ8 {
  int8_t* k;
  int8_t* v;
  float* k_scale;
  float* v_scale;
  __nv_bfloat16* v_bf16;
  int ctxp;   // context rounded up to a multiple of 4
};
void launch_rope_scatter_q8(const float* qkv, const PartialSums& ps, const RingState& rs, int B, int H, int D, int context,
                            const float* inv_freq, __nv_bfloat16* q_out, const KvRingQ8& ring, cudaStream_t st);
void launch_lm_attention_q8(const __nv_bfloat16* q, const RingState& rs, const uint8_t* mask, int B, int H, int D, int context,
                            const KvRingQ8& ring, __nv_bfloat16* out, cudaStream_t st);
// h f32 [B, 2*Hd] -> g bf16 [B, Hd] = silu(h[:, :Hd]) * h[:, Hd:]   (batched_transformer.rs:170-176)
void launch_silu_gate(const float* h, int B, int Hd, __nv_bfloat16* g, cudaStream_t st);
// greedy sampling (asr.rs:208-210): first argmax over logits [B, V]; partial planes of the text head are summed into
// `logits` on the way
void launch_argmax(float* logits, const PartialSums& ps, int B, int V, uint32_t* tokens, cudaStream_t st);
// extra heads (lm.rs:948-955 + asr.rs:195-206): prs[h*B + b] = softmax(W_h . ys[b])[0]
void launch_extra_heads(const __nv_bfloat16* ys, const __nv_bfloat16* w /*[nh][dim][D]*/, int nh, int dim, int B, int D,
                        float* prs, cudaStream_t st);
// token feed (asr.rs:133-145,172-176): lm_audio_in = first_step[b] ? pad : next_cb[b]; lm_text_in =
// first_step[b] ? start : text_tok[b]
void launch_asr_prepare(const uint32_t* next_cb, const uint8_t* first_step, const uint32_t* text_tok, uint32_t pad,
                        uint32_t start, int B, int n_cb, uint32_t* lm_audio_in, uint32_t* lm_text_in, cudaStream_t st);
// state update for active streams (asr.rs:177-183,221-226): next_cb <- codes, text_tok <- sampled, first_step <- 0
void launch_asr_commit(const uint32_t* codes, const uint8_t* mask, const uint32_t* sampled, int B, int n_cb,
                       uint32_t* next_cb, uint32_t* text_tok, uint8_t* first_step, cudaStream_t st);

}  // namespace dsm
