// Mimi handle: streaming SEANet encoder/decoder, codec transformers, split RVQ.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "gemm_simt.cuh"
#include "gemm_tc6.cuh"
#include "kernels.cuh"
#include "runtime.cuh"

namespace dsm {

// One codec GEMM: runs on the tcgen05 split-bf16 kernel (f32-accurate) or on the CUDA-core kernel
// (DSM_CODEC_GEMM=simt, or shapes the tensor-core kernel does not take).
struct CodecGemm {
  GemmParams g;      // g.W = f32 weights [N, K] on the device
  Tc6Gemm t;
  bool use_tc = false;
  const __nv_bfloat16* w_planes = nullptr;   // kept so that the GEMM can be re-tuned for another SM budget
  Tc6Workspace wsp;
};
// Every codec GEMM built while a registry is installed (MimiImpl's constructor) is recorded in it.
struct CodecGemmRegistry {
  std::vector<std::shared_ptr<CodecGemm>> all;
};
void codec_gemm_set_registry(CodecGemmRegistry* r);
void codec_gemms_retune(CodecGemmRegistry& r, int sms);
bool codec_gemm_tc_enabled();
bool codec_presplit_enabled();
std::shared_ptr<CodecGemm> make_codec_gemm(Arena& arena, const GemmParams& g, const std::vector<float>& w_host,
                                           const Tc6Workspace& wsp = Tc6Workspace());
void run_codec_gemm(const CodecGemm& c, cudaStream_t st);

// One f32 streaming transformer with per-stream ring KV (batched_transformer.rs, Mimi flavour:
// LayerNorm, GELU MLP, LayerScale, RoPE).
struct CodecTransformer {
  int B = 0, T = 2, D = 512, H = 8, hd = 64, ff = 2048, ctx = 250, L = 8;
  float eps = 1e-5f;
  RingState ring{};
  float* inv_freq = nullptr;
  float *xn = nullptr, *qkv = nullptr, *attn = nullptr, *hbuf = nullptr;
  // the same activations as three bf16 planes ([3][rows][cols]) for the pre-split GEMM path (null when off)
  __nv_bfloat16 *xn_pl = nullptr, *attn_pl = nullptr, *h_pl = nullptr;
  struct Layer {
    float *ln1w, *ln1b, *ln2w, *ln2b, *ls1, *ls2;
    float *k_ring, *v_ring;
    std::vector<float> h_in_proj, h_out_proj, h_lin1, h_lin2;   // host copies until the plan is built
  };
  Arena* arena_ = nullptr;
  Tc6Workspace wsp;
  std::vector<Layer> layers;
  size_t kv_bytes = 0;

  void init(Arena& arena, const dsm_mimi_spec& s, int batch, const std::string& prefix,
            const std::function<std::vector<float>(const std::string&, size_t, float, float)>& get);
  // x rows (b, t) addressed through x_map; updated in place.
  void append_plan(Plan& plan, float* x, RowMap x_map, const uint8_t* d_mask);
  void reset_all(cudaStream_t st);
  void reset_batch_idx(int b, cudaStream_t st);
};

struct SeanetLayer {
  std::string wname;   // weight prefix (without the trailing .weight / .bias)
  int cin, cout, k, stride;
  bool elu_in;         // ELU applied to this layer's input
  int res;             // 0 none, 1 first conv of a resnet block, 2 last conv (adds the block input)
  bool tr;             // conv-transpose
};

struct MimiImpl {
  dsm_mimi_spec spec{};
  int B = 0, device = 0;
  cudaStream_t stream = nullptr;
  Arena arena;
  uint8_t* d_mask = nullptr;
  uint8_t* h_mask = nullptr;       // pinned staging
  uint32_t* h_codes = nullptr;     // pinned staging
  int* d_first_step = nullptr;     // 1 until the first encode step ran (left_pad_applied, conv.rs:232)
  int* d_first_step_dec = nullptr;
  Tc6Workspace wsp;                // split-K scratch shared by the handle's codec GEMMs
  CodecGemmRegistry gemms;         // every codec GEMM of both plans (re-tuned when the handle moves to an SM partition)
  std::vector<cudaStream_t> retired_streams;
  int sms = 148;                   // SMs the plans are tuned for
  // An ASR state that partitions the SMs between its branches (runtime.cuh) hands the codec its stream and budget.
  void move_to_partition(cudaStream_t st, int sm_budget);

  // ---- encoder ----
  float* d_pcm = nullptr;          // [B, 1920]
  float* enc_state0 = nullptr;     // [B, k0-1] left context of the C_in=1 conv
  float* x_ds = nullptr;           // [B, 2+2, 512]: downsample conv input; rows [2,4) = transformer stream
  float* dbg_seanet = nullptr;     // [B, 2, 512] copy of the SEANet output (parity taps)
  float* dbg_xfmr = nullptr;       // [B, 2, 512] copy of the codec transformer output
  float* latent = nullptr;         // [B, 512]
  float* resid = nullptr;          // [B, 512] = [first-proj | rest-proj] residuals
  float* scores = nullptr;         // [B, bins]
  uint32_t* d_codes = nullptr;     // [B, n_q]
  float* emb_all = nullptr;        // [n_q][bins][qdim] codebook embeddings (quantization.rs:91-93)
  float* c2_all = nullptr;         // [n_q][bins]  |e|^2 / 2 (quantization.rs:95)
  CodecTransformer enc_tr, dec_tr;
  std::vector<CommitEntry> enc_commits, dec_commits;
  CommitEntry *d_enc_commits = nullptr, *d_dec_commits = nullptr;
  int enc_max_count = 0, dec_max_count = 0;
  Plan enc_plan, dec_plan;
  GraphExec enc_graph, dec_graph;

  // ---- decoder ----
  uint32_t* d_codes_in = nullptr;  // [B, n_q]
  float* q_sum = nullptr;          // [B, 512] = [sum of first | sum of rest] codebook rows
  float* dec_emb = nullptr;        // [B, 512] quantiser output
  float* up_state = nullptr;       // [B, 512] previous dec_emb (depthwise upsample carry)
  float* d_pcm_out = nullptr;      // [B, 1920]
  float* h_pcm_out = nullptr;

  MimiImpl(const dsm_mimi_spec& spec, const dsm_weights* w, int device, int batch);
  ~MimiImpl();

  void upload_mask(const uint8_t* host_mask, cudaStream_t st);
  // Enqueue the encode of d_pcm under `mask_dev` on `st`; codes land in d_codes.
  void enqueue_encode(const uint8_t* mask_dev, cudaStream_t st);
  void enqueue_decode(const uint8_t* mask_dev, cudaStream_t st);
  void encode_step(const float* pcm, const uint8_t* mask, uint32_t* codes);
  void decode_step(const uint32_t* codes, const uint8_t* mask, float* pcm);
  void reset_state();
  void reset_batch_idx(int b);

 private:
  void build_encoder(const std::function<std::vector<float>(const std::string&, size_t, float, float)>& get,
                     const WeightTable& wt);
  void build_decoder(const std::function<std::vector<float>(const std::string&, size_t, float, float)>& get,
                     const WeightTable& wt);
};

}  // namespace dsm
