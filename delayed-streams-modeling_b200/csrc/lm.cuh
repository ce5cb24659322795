// LM handle: delayed-streams transformer step (lm.rs LmModel without depformer) and the ASR state.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "runtime.cuh"

namespace dsm {

struct MimiImpl;

struct LmImpl {
  dsm_lm_spec spec{};
  int B = 0, device = 0;
  int D = 0, H = 0, hd = 0, hidden = 0, L = 0, ctx = 0;
  cudaStream_t stream = nullptr;
  Arena arena;
  uint8_t* d_mask = nullptr;
  RingState ring{};
  float* inv_freq = nullptr;

  // weights (bf16 matrices, f32 norm scales)
  __nv_bfloat16* text_emb = nullptr;    // [text_in_vocab, D]
  __nv_bfloat16* audio_emb = nullptr;   // [n_cb][audio_vocab][D]
  struct Layer {
    __nv_bfloat16 *in_proj, *out_proj, *lin_in, *lin_out;
    float *norm1, *norm2;
    __nv_bfloat16 *k_ring, *v_ring;     // [B, H, ctx, hd]   (kv_dtype == DSM_KV_BF16)
    KvRingQ8 q8;                        // int8 rings + scales  (kv_dtype == DSM_KV_INT8)
    TcGemm g_qkv, g_out, g_in, g_down;  // tcgen05 GEMM descriptors (tensor maps over the buffers below)
  };
  TcGemm g_text;
  bool use_tc = true;                   // DSM_LM_GEMM=simt selects the CUDA-core GEMM (A/B debugging)
  std::vector<Layer> layers;
  float* out_norm = nullptr;
  __nv_bfloat16* text_linear = nullptr; // [V, D]
  __nv_bfloat16* extra_w = nullptr;     // [nh][dim][D]
  size_t weight_bytes = 0, kv_bytes = 0;

  // per-step buffers
  uint32_t* d_text_ids = nullptr;       // [B]
  uint32_t* d_audio_ids = nullptr;      // [B, n_cb]
  float* d_cond = nullptr;              // [D]
  int has_cond = 0;
  float* x = nullptr;                   // [B, D] residual stream (f32)
  __nv_bfloat16* xn = nullptr;          // [B, D]
  float* qkv = nullptr;                 // [B, 3D]
  __nv_bfloat16* q = nullptr;           // [B, D]
  __nv_bfloat16* attn = nullptr;        // [B, D]
  float* hbuf = nullptr;                // [B, 2*hidden]
  __nv_bfloat16* gbuf = nullptr;        // [B, hidden]
  __nv_bfloat16* ys_bf16 = nullptr;     // [B, D]
  float* ys_f32 = nullptr;              // [B, D]
  float* logits = nullptr;              // [B, V]
  uint32_t* d_sampled = nullptr;        // [B] greedy argmax of the logits (last kernel of the plan)

  // optional per-launch timing of the attention kernel (bench.py's roofline leg; eager mode only)
  bool profile_attn = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  size_t prof_used = 0;
  void profile_begin(cudaStream_t st);
  void profile_end(cudaStream_t st);
  void profile_read(double* ms_sum, long long* launches);

  Plan plan;        // forward without conditioning
  Plan plan_cond;   // forward with Condition::AddToInput
  GraphExec graph, graph_cond;

  int kv_dtype = DSM_KV_BF16;
  LmImpl(const dsm_lm_spec& spec, const dsm_weights* w, int device, int batch, const dsm_lm_options* opt = nullptr);
  ~LmImpl();
  void upload_mask(const uint8_t* host_mask, cudaStream_t st);
  void forward(const uint32_t* text_ids, const uint32_t* audio_ids, const float* cond, const uint8_t* mask,
               float* logits_out, float* ys_out);
  void reset_state();
  void reset_batch_idx(int b);
  // An ASR state that partitions the SMs between its branches (runtime.cuh) hands the LM its stream and budget.
  void move_to_partition(cudaStream_t st, int sm_budget);
  int sms = 148;                       // SMs the plans are tuned for

 private:
  void build_plan(Plan& p, bool with_cond);
  void tune_plans(int sm_budget);
  float* partial_buf = nullptr;
  size_t partial_buf_bytes = 0;
  std::vector<cudaStream_t> retired_streams;
};

struct ItemState {   // asr.rs:16-51
  int step_idx = 0;
  std::vector<uint32_t> word_tokens;
  bool unended_word = false;
  double last_stop_time = 0.0;
};

struct AsrMsgStore {
  int kind, batch_idx, step_idx;
  std::vector<uint32_t> tokens;
  double time;
};

struct AsrImpl {
  MimiImpl* mimi = nullptr;
  LmImpl* lm = nullptr;
  int B = 0, n_cb = 0, delay = 0, nh = 0;
  int model_step_idx = 0;
  Arena arena;
  uint32_t* next_cb = nullptr;      // [B, n_cb] codes of the previous frame (asr.rs:61)
  uint32_t* text_tok = nullptr;     // [B] last sampled text token
  uint8_t* first_step = nullptr;    // [B] ItemState::is_first_step
  uint32_t* codes_in = nullptr;     // [B, n_cb] codes of this frame when supplied by the caller
  uint32_t* sampled = nullptr;      // [B]
  float* prs = nullptr;             // [nh, B]
  uint32_t* h_sampled = nullptr;    // pinned
  float* h_prs = nullptr;           // pinned
  std::vector<uint8_t> h_mask;
  std::vector<ItemState> items;
  std::vector<AsrMsgStore> msgs;
  std::vector<float> prs_copy;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_m0 = nullptr, ev_m1 = nullptr,
              ev_l0 = nullptr, ev_l1 = nullptr;
  GraphExec graph_pcm, graph_tokens;
  int sm_split_lm = 0, sm_split_codec = 0;   // SMs of the two branch partitions (0 = the branches share all SMs)
  float last_ms[3] = {-1.f, -1.f, -1.f};

  AsrImpl(MimiImpl* mimi, LmImpl* lm, int delay, float temperature);
  ~AsrImpl();
  void step(const float* pcm, const uint32_t* codes, const uint8_t* mask, uint32_t* text_tokens);
  void step_begin(const float* pcm, const uint32_t* codes, const uint8_t* mask);   // asynchronous half
  void step_end(uint32_t* text_tokens);                                            // waits, runs the word machine
  bool in_flight = false, flight_graph = false, flight_mimi = false;
  void reset();
  void reset_batch_idx(int b);

 private:
  void enqueue(bool with_mimi, bool timing);
  void host_update(const uint8_t* mask);
};

}  // namespace dsm
