/*
 * dsm_b200.h -- C ABI of the B200-native batched streaming-STT hot path.
 *
 * This header is the drop-in boundary.  Each entry point replaces one method of the reference's Rust
 * `moshi` crate (paths relative to /root/reference/server/rust/moshi/moshi-core/src unless noted);
 * the citation next to each declaration names the method it stands in for.  The reference has no
 * FFI of its own (everything below the Rust methods is candle), so a maintainer binds these with a
 * plain `extern "C"` block -- see INTEGRATION.md for the Rust shim.
 *
 * Conventions
 *   - Every function returns a dsm_status (0 = ok, < 0 = error); the message of the last error on the
 *     calling thread is available through dsm_last_error().  Nothing aborts or throws across the ABI
 *     (the reference returns candle::Result and bails, e.g. asr.rs:158-160).
 *   - Handles own all device memory (weights, conv/KV state, workspaces) and one CUDA stream each.
 *     They are thread-compatible, not thread-safe (the reference's objects are `&mut self`).
 *   - Data pointers are plain row-major, batch-major buffers.  Unless stated otherwise a data pointer
 *     may be host memory (pinned or pageable) or device memory; the library detects which.
 *   - `mask` is `[B]` bytes on the HOST, 1 = stream active this step (StreamMask::new,
 *     streaming.rs:42-47); NULL means all active.
 *   - There is no CPU fallback: creating a handle without a CUDA device fails with DSM_ERR_CUDA.
 */
#ifndef DSM_B200_H
#define DSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t dsm_status;
enum {
  DSM_OK = 0,
  DSM_ERR_INVALID_ARG = -1,
  DSM_ERR_SHAPE = -2,
  DSM_ERR_CUDA = -3,
  DSM_ERR_OOM = -4,
  DSM_ERR_UNSUPPORTED = -5,
  DSM_ERR_MISSING_WEIGHT = -6,
  DSM_ERR_CAPACITY = -7 /* no free slot (the reference closes the socket with code 4000, protocol.rs:33) */
};

enum { DSM_F32 = 0, DSM_BF16 = 1, DSM_F16 = 2, DSM_U32 = 3, DSM_I64 = 4 /* token logs only */ };

#define DSM_FRAME_SIZE 1920 /* moshi-server/src/batched_asr.rs:26 */

/* One named host tensor (what candle's VarBuilder hands the reference's constructors). */
typedef struct dsm_tensor {
  const char* name;
  int32_t dtype; /* DSM_F32 / DSM_BF16 / DSM_F16 */
  int32_t ndim;
  int64_t shape[4];
  const void* data; /* host pointer, contiguous */
} dsm_tensor;

typedef struct dsm_weights {
  const dsm_tensor* tensors;
  int64_t count;
} dsm_weights;

/* mimi::Config::v0_1 (mimi.rs:32-93).  Zero-initialise then call dsm_mimi_spec_v0_1(). */
typedef struct dsm_mimi_spec {
  int32_t n_q;             /* quantizer_n_q; 32 for STT/TTS (batched_asr.rs:754) */
  int32_t bins;            /* 2048 */
  int32_t qdim;            /* 256 */
  int32_t dim;             /* 512: seanet dimension == transformer d_model */
  int32_t n_filters;       /* 64 */
  int32_t ratios[4];       /* {8,6,5,4}; the encoder walks them reversed */
  int32_t kernel_size;     /* 7 */
  int32_t residual_kernel_size; /* 3 */
  int32_t last_kernel_size;     /* 3 */
  int32_t compress;        /* 2 */
  int32_t num_heads;       /* 8 */
  int32_t num_layers;      /* 8 */
  int32_t dim_feedforward; /* 2048 */
  int32_t context;         /* 250 */
  float max_period;        /* 10000 */
  float ln_eps;            /* 1e-5 */
  int32_t downsample_stride; /* 2 */
  int32_t with_encoder;    /* build encode_step state */
  int32_t with_decoder;    /* build decode_step state */
} dsm_mimi_spec;

/* lm::Config as the STT TOMLs fill it (configs/stt/config-stt-en_fr-hf.toml:27-56). */
typedef struct dsm_lm_spec {
  int32_t d_model;
  int32_t num_heads;
  int32_t num_layers;
  int32_t dim_feedforward;
  int32_t context;
  float max_period;
  int32_t text_in_vocab_size;
  int32_t text_out_vocab_size;
  int32_t audio_vocab_size;
  int32_t audio_codebooks;
  int32_t extra_heads_num;
  int32_t extra_heads_dim;
  float rms_eps; /* 1e-8, batched_transformer.rs:246-248 */
} dsm_lm_spec;

typedef struct dsm_mimi dsm_mimi;
typedef struct dsm_lm dsm_lm;
typedef struct dsm_asr dsm_asr;

/* ---- library ---------------------------------------------------------------------------- */
const char* dsm_version(void);
/* Copies the calling thread's last error message (NUL-terminated) into buf; returns its length. */
size_t dsm_last_error(char* buf, size_t cap);
/* Number of CUDA devices visible (0 when none: every create call will then fail). */
int32_t dsm_device_count(void);
/* Kernel launches issued by this library on the calling thread (graph replays count their kernels). */
/* 1 (default): steps replay a captured CUDA graph; 0: eager launches (debugging / per-kernel timing).
 * The environment variable DSM_NO_GRAPH=1 has the same effect as dsm_set_graphs(0). */
void dsm_set_graphs(int32_t enabled);
long long dsm_launch_count(void);
void dsm_launch_count_reset(void);

/* Stand-alone check of the projection GEMM kernels on device 0: out = X[M,K] . W[N,K]^T with bf16 bit
 * patterns in and f32 accumulation, through the CUDA-core kernel (impl 0) or the tcgen05 kernel
 * (impl 1).  mode 0: store f32 [M,N]; 1: accumulate into inout_f32 [M,N]; 2: gated pair ->
 * out_bf16 [M,N/2] = silu(acc[:, j]) * acc[:, N/2 + j] (batched_transformer.rs:170-176). */
dsm_status dsm_debug_gemm_bf16(int32_t impl, int32_t mode, int32_t M, int32_t N, int32_t K, const uint16_t* X,
                               const uint16_t* W, float* inout_f32, uint16_t* out_bf16);

/* Stand-alone check of the codec GEMM kernels on device 0: out[(b,t), n] = bias[n] + sum_k
 * A[b*stride_b + t*stride_t + k] * W[n,k] (the implicit-GEMM streaming-conv window, conv.rs:342), f32,
 * through the CUDA-core kernel (impl 0) or the tcgen05 3-way-bf16-split kernel (impl 1). */
dsm_status dsm_debug_gemm_f32(int32_t impl, int32_t B, int32_t T, int64_t stride_b, int64_t stride_t, int32_t N, int32_t K,
                              const float* A, int64_t a_elems, const float* W, const float* bias, float* out);

/* Known-answer access to the DEVICE ring bookkeeping (ScatteredCacheBuilder::indices_and_mask, kv_cache.rs:119-237,
 * test table :338-405): a fresh ring of `batch` streams and `context` slots is stepped n_steps times on device 0; step i
 * appends t[i] tokens under masks[i*batch .. ] (1 = active).  Outputs, concatenated over the steps: the scatter indices
 * [batch, t[i]] u32 and the additive mask [batch, t[i], context] f32 (0 / -inf) exactly as the attention kernels
 * reconstruct visibility (the same device functions). */
dsm_status dsm_debug_ring_trace(int32_t batch, int32_t context, int32_t n_steps, const int32_t* t, const uint8_t* masks,
                                uint32_t* indices_out, float* mask_out);

void dsm_mimi_spec_v0_1(dsm_mimi_spec* spec, int32_t n_q);            /* mimi.rs:32 Config::v0_1 */
void dsm_lm_spec_stt_1b_en_fr(dsm_lm_spec* spec);                     /* config-stt-en_fr-hf.toml */
void dsm_lm_spec_stt_2_6b_en(dsm_lm_spec* spec);                      /* config-stt-en-hf.toml */

/* ---- Mimi ------------------------------------------------------------------------------- */
/* Mimi::batched (mimi.rs:113-171).  Folds weight-norm (conv.rs:27-45) and derives codebook
 * embeddings / c2 (quantization.rs:86-107) at load.  weights == NULL => seeded device-side random
 * init (benchmarks only). */
dsm_status dsm_mimi_create(const dsm_mimi_spec* spec, const dsm_weights* weights, int32_t device,
                           int32_t batch_size, dsm_mimi** out);
void dsm_mimi_destroy(dsm_mimi* h);
/* Mimi::encode_step (mimi.rs:195-206).  pcm [B,1920] f32 -> codes [B,n_q] u32; *produced = 1 when a
 * frame of codes was written (always, for 1920-sample steps).  Asynchronous on the handle's stream
 * when `codes` is device memory; synchronous when it is host memory. */
dsm_status dsm_mimi_encode_step(dsm_mimi* h, const float* pcm, const uint8_t* mask, uint32_t* codes,
                                int32_t* produced);
/* Mimi::decode_step (mimi.rs:217-225).  codes [B,n_q] u32 -> pcm [B,1920] f32. */
dsm_status dsm_mimi_decode_step(dsm_mimi* h, const uint32_t* codes, const uint8_t* mask, float* pcm,
                                int32_t* produced);
/* Latent that the last encode_step fed to the quantiser, [B,dim] f32 (Mimi::encode_pre_quantize's
 * streaming counterpart, mimi.rs:177-184); for parity checks. */
dsm_status dsm_mimi_last_latent(dsm_mimi* h, float* latent);
/* Parity taps: "seanet_out" [B,2,dim], "transformer_out" [B,2,dim], "latent" [B,dim] (encoder side,
 * values of the last encode_step, time-major rows), "dec_emb" [B,dim] (quantiser output of the last
 * decode_step).  out == NULL just reports *numel. */
dsm_status dsm_mimi_debug_tensor(dsm_mimi* h, const char* name, float* out, int64_t cap, int64_t* numel);
dsm_status dsm_mimi_reset_state(dsm_mimi* h);                          /* mimi.rs:227-234 */
dsm_status dsm_mimi_reset_batch_idx(dsm_mimi* h, int32_t batch_idx);   /* mimi.rs:236-244 */
dsm_status dsm_mimi_synchronize(dsm_mimi* h);
int32_t dsm_mimi_batch_size(const dsm_mimi* h);
int32_t dsm_mimi_n_q(const dsm_mimi* h);
int64_t dsm_mimi_device_bytes(const dsm_mimi* h);
int64_t dsm_mimi_kv_bytes(const dsm_mimi* h);

/* ---- LM --------------------------------------------------------------------------------- */
/* LmModel::batched (lm.rs:812-897), depformer-less (STT).  Weights are rounded to bf16 on load
 * (model_dtype == bf16 on CUDA, moshi-server/src/utils.rs:386-395). */
dsm_status dsm_lm_create(const dsm_lm_spec* spec, const dsm_weights* weights, int32_t device,
                         int32_t batch_size, dsm_lm** out);
/* Storage format of the LM's ring KV cache (kv_cache.rs:20-51 keeps it in the model dtype = bf16, the default here).
 * DSM_KV_INT8 halves the ring (98.3 -> 50.8 MB per stt-1b stream, the route past ~1,700 streams per 180 GB GPU):
 * int8 K / V with one f32 scale per (stream, head, slot), dequantised inside the attention kernel.  It is an
 * approximation of the reference (measured logits error in DESIGN.md section 4), hence opt-in.
 * DSM_KV_INT8K keeps V in bf16 and only K in int8 (74.5 MB per stream): three quarters of the bytes, half the error. */
enum { DSM_KV_BF16 = 0, DSM_KV_INT8 = 1, DSM_KV_INT8K = 2 };
typedef struct dsm_lm_options {
  int32_t kv_dtype;     /* DSM_KV_* */
  int32_t reserved[7];  /* zero */
} dsm_lm_options;
/* dsm_lm_create with options (NULL = defaults = dsm_lm_create). */
dsm_status dsm_lm_create_ex(const dsm_lm_spec* spec, const dsm_weights* weights, int32_t device, int32_t batch_size,
                            const dsm_lm_options* options, dsm_lm** out);
int32_t dsm_lm_kv_dtype(const dsm_lm* h);
void dsm_lm_destroy(dsm_lm* h);
/* LmModel::forward_cond (lm.rs:957-1008): text_ids [B] u32, audio_ids [B,n_cb] u32,
 * cond [d_model] f32 or NULL (Condition::AddToInput) -> logits [B,V] f32 and ys [B,D] f32 (either
 * may be NULL).  The reference returns bf16 tensors; f32 here is the same values widened. */
dsm_status dsm_lm_forward(dsm_lm* h, const uint32_t* text_ids, const uint32_t* audio_ids,
                          const float* cond, const uint8_t* mask, float* logits, float* ys);
dsm_status dsm_lm_reset_state(dsm_lm* h);                              /* lm.rs:899-902 */
dsm_status dsm_lm_reset_batch_idx(dsm_lm* h, int32_t batch_idx);       /* lm.rs:1108-1110 */
dsm_status dsm_lm_synchronize(dsm_lm* h);
int32_t dsm_lm_text_start_token(const dsm_lm* h);                      /* lm.rs:916-918 */
int32_t dsm_lm_audio_pad_token(const dsm_lm* h);                       /* lm.rs:912-914 */
int32_t dsm_lm_in_audio_codebooks(const dsm_lm* h);                    /* lm.rs:904-906 */
/* Bytes of device memory the handle's ring KV cache occupies (for the roofline arithmetic). */
/* The handle's cudaStream_t (for recording caller-side CUDA events around steps). */
void* dsm_lm_stream(const dsm_lm* h);
void* dsm_mimi_stream(const dsm_mimi* h);
/* Per-launch CUDA-event timing of the ring-attention kernel (eager mode only; measurement aid):
 * enable, run steps, then read the summed kernel time and the number of launches timed. */
dsm_status dsm_lm_profile_attention(dsm_lm* h, int32_t enable);
dsm_status dsm_lm_profile_read(dsm_lm* h, double* ms_sum, long long* launches);
/* Profiling aid: pretend every stream has already consumed `pos` frames (ring contents are left as
 * they are), so a short profiled run reads full rings without replaying `context` steps first. */
dsm_status dsm_lm_debug_set_position(dsm_lm* h, uint32_t pos);
int64_t dsm_lm_kv_bytes(const dsm_lm* h);
int64_t dsm_lm_weight_bytes(const dsm_lm* h);

/* ---- ASR state -------------------------------------------------------------------------- */
/* asr::AsrMsg (asr.rs:9-13). */
enum { DSM_MSG_STEP = 0, DSM_MSG_WORD = 1, DSM_MSG_END_WORD = 2 };
typedef struct dsm_asr_msg {
  int32_t kind;
  int32_t batch_idx;      /* WORD / END_WORD */
  int32_t step_idx;       /* STEP: model_step_idx after the step */
  int32_t n_tokens;       /* WORD */
  const uint32_t* tokens; /* WORD: valid until the next step call on this handle */
  double time;            /* WORD: start_time, END_WORD: stop_time (seconds) */
  const float* prs;       /* STEP: [n_heads * B], prs[h*B + b]; valid until the next step */
} dsm_asr_msg;

/* asr::State::new (asr.rs:64-88).  Borrows mimi and lm (they must outlive the state; mimi may be
 * NULL when only step_tokens is used).  temperature must be <= 0 (greedy, asr.rs:208-210). */
dsm_status dsm_asr_create(dsm_mimi* mimi, dsm_lm* lm, int32_t asr_delay_in_tokens, float temperature,
                          dsm_asr** out);
void dsm_asr_destroy(dsm_asr* h);
/* asr::State::step_pcm (asr.rs:115-131): Mimi encode of this frame runs concurrently with the LM
 * step on the previous frame's codes (asr.rs:172-183).  text_tokens [B] u32 HOST out (may be NULL).
 * Messages of the step are read with dsm_asr_msg_count / dsm_asr_msg_get. */
dsm_status dsm_asr_step_pcm(dsm_asr* h, const float* pcm, const uint8_t* mask, uint32_t* text_tokens);
/* asr::State::step_tokens / step_tokens_vec (asr.rs:147-255, 268-283): codes [B,n_cb] u32. */
dsm_status dsm_asr_step_tokens(dsm_asr* h, const uint32_t* codes, const uint8_t* mask,
                               uint32_t* text_tokens);
int32_t dsm_asr_msg_count(const dsm_asr* h);
dsm_status dsm_asr_msg_get(const dsm_asr* h, int32_t i, dsm_asr_msg* out);
dsm_status dsm_asr_reset(dsm_asr* h);                                  /* asr.rs:104-113 */
dsm_status dsm_asr_reset_batch_idx(dsm_asr* h, int32_t batch_idx);     /* asr.rs:257-266 */
int32_t dsm_asr_batch_size(const dsm_asr* h);                          /* asr.rs:94-96 */
int32_t dsm_asr_model_step_idx(const dsm_asr* h);                      /* asr.rs:90-92 */
int32_t dsm_asr_delay_in_tokens(const dsm_asr* h);                     /* asr.rs:98-100 */
/* Text logits [B,V] f32 of the last step (what asr.rs:208-210 takes the argmax of); parity only. */
dsm_status dsm_asr_last_logits(dsm_asr* h, float* logits);
/* Device-side timing of the last step (ms) measured with CUDA events on the handle's streams:
 * out[0] = whole step, out[1] = Mimi encode branch, out[2] = LM branch. */
dsm_status dsm_asr_last_step_ms(dsm_asr* h, float out[3]);
/* SM partition of the state's two branches (no reference counterpart: candle runs one stream).  A state built from
 * a Mimi encoder and an LM confines the LM branch (HBM-bound ring attention) and the codec branch (tensor / FMA-bound)
 * to disjoint SM sets -- CUDA green contexts -- when that was measured to pay (bf16 ring: always; 8-bit rings: up to
 * 1,024 streams; DSM_SM_SPLIT=<LM SMs>
 * forces a size, 0 disables).  out[0] = SMs of the LM branch, out[1] = SMs of the codec branch; both 0 when the
 * branches share the device.  The handles passed to dsm_asr_create then run on the partition's streams (dsm_lm_stream
 * / dsm_mimi_stream return the new ones; handles obtained earlier stay valid but idle). */
dsm_status dsm_asr_sm_partition(const dsm_asr* h, int32_t out[2]);

/* ---- batched step scheduler ---------------------------------------------------------------- */
/* Host side of moshi-server's BatchedAsr (moshi-server/src/batched_asr.rs:216-717) without the network
 * front end: a fixed pool of B slots with a PCM queue each; one tick gathers at most one frame per
 * active slot into a pinned [B,1920] buffer, builds the mask, applies pending per-slot resets, runs
 * dsm_asr_step_pcm and fans the results out per slot.  push_* may be called from other threads while a
 * tick runs. */
typedef struct dsm_sched dsm_sched;
enum { DSM_SCHED_STEP = 0, DSM_SCHED_WORD = 1, DSM_SCHED_END_WORD = 2, DSM_SCHED_MARKER = 3, DSM_SCHED_READY = 4 };
/* OutMsg (moshi-server/src/asr.rs:24-34) before detokenisation. */
typedef struct dsm_sched_out {
  int32_t kind;
  int32_t slot;
  uint64_t channel_id;    /* connection the message belongs to (a slot is re-used across connections) */
  int32_t step_idx;
  int32_t n_tokens;       /* WORD */
  const uint32_t* tokens; /* WORD: valid until the next tick */
  double time;            /* WORD start_time / END_WORD stop_time */
  int32_t n_prs;          /* STEP */
  const float* prs;       /* STEP: this slot's VAD head probabilities */
  int64_t marker_id;      /* MARKER */
  int64_t buffered_pcm;   /* STEP / READY: samples still queued for the slot */
} dsm_sched_out;
/* asr may be NULL for a dry scheduler (slot / queue / marker logic only, no device step); then
 * batch_size_if_no_asr gives B. */
dsm_status dsm_sched_create(dsm_asr* asr, int32_t batch_size_if_no_asr, dsm_sched** out);
/* One scheduler over n_devices GPUs: asrs[g] is the asr::State living on device g (all with the same batch B);
 * device g owns slots [g*B, (g+1)*B) -- streams never interact, so routing is device = slot / B and there is no
 * step-path exchange.  asrs == NULL (or all NULL): dry scheduler over n_devices x batch_size_if_no_asr slots. */
dsm_status dsm_sched_create_multi(dsm_asr* const* asrs, int32_t n_devices, int32_t batch_size_if_no_asr, dsm_sched** out);
int32_t dsm_sched_num_devices(const dsm_sched* h);
int32_t dsm_sched_slots_per_device(const dsm_sched* h);
void dsm_sched_destroy(dsm_sched* h);
/* BatchedAsr::channels (batched_asr.rs:795-808): takes a free slot or fails with DSM_ERR_CAPACITY.  The
 * slot's model state is reset before its first step (InMsg::Init, :567-573). */
dsm_status dsm_sched_open(dsm_sched* h, int32_t* slot, uint64_t* channel_id);
dsm_status dsm_sched_close(dsm_sched* h, int32_t slot);                                   /* :633-652 */
dsm_status dsm_sched_push_audio(dsm_sched* h, int32_t slot, const float* pcm, int64_t n); /* InMsg::Audio */
dsm_status dsm_sched_push_marker(dsm_sched* h, int32_t slot, int64_t id);                 /* InMsg::Marker */
/* One pass of encoder_loop + model_loop + post_process (:314-522, 655-717).  *stepped = 0 when no slot had
 * a full frame (the reference then sleeps 2 ms, :398-400). */
dsm_status dsm_sched_tick(dsm_sched* h, int32_t* stepped);
int32_t dsm_sched_out_count(const dsm_sched* h);
dsm_status dsm_sched_out_get(const dsm_sched* h, int32_t i, dsm_sched_out* out);
/* Variants that name the connection: a stale client whose slot was closed, reclaimed and re-opened by someone else is
 * refused (DSM_ERR_INVALID_ARG) instead of reaching the new connection (the reference's per-connection channels,
 * batched_asr.rs:110-118).  channel_id 0 = unchecked.  A slot has ONE producer thread (its connection): the PCM and
 * marker queues are lock-free single-producer / single-consumer rings (2^17 samples = 5.5 s; DSM_ERR_CAPACITY when a
 * push does not fit). */
dsm_status dsm_sched_push_audio_ch(dsm_sched* h, int32_t slot, uint64_t channel_id, const float* pcm, int64_t n);
dsm_status dsm_sched_push_marker_ch(dsm_sched* h, int32_t slot, uint64_t channel_id, int64_t id);
dsm_status dsm_sched_close_ch(dsm_sched* h, int32_t slot, uint64_t channel_id);
/* Threaded mode: one thread per device runs the pipelined loop (gather of frame t+1 into the other half of a
 * double-buffered pinned batch while the GPU runs frame t; sleeps 2 ms when no slot has a frame, batched_asr.rs:398-400)
 * and queues each connection's messages; the connection's thread drains them with dsm_sched_poll.  dsm_sched_tick is
 * refused while the threads run. */
typedef struct dsm_sched_event {
  int32_t kind, slot, step_idx, n_tokens, n_prs;
  uint64_t channel_id;
  double time;
  int64_t marker_id, buffered_pcm;
  uint32_t tokens[64];
  float prs[8];
} dsm_sched_event;
dsm_status dsm_sched_start(dsm_sched* h);
dsm_status dsm_sched_stop(dsm_sched* h);
dsm_status dsm_sched_poll(dsm_sched* h, int32_t slot, dsm_sched_event* ev, int32_t* got);
typedef struct dsm_sched_stats {
  int64_t ticks, steps, idle_sleeps, frames;
  int64_t gather_ns, step_wait_ns, fanout_ns;   /* summed over the ticks of the device */
} dsm_sched_stats;
dsm_status dsm_sched_get_stats(const dsm_sched* h, int32_t device, dsm_sched_stats* out);
void dsm_sched_set_dry_step_us(dsm_sched* h, int32_t us);   /* dry mode: pretend a device step takes this long */
int64_t dsm_sched_device_step_idx(const dsm_sched* h, int32_t device);
int32_t dsm_sched_free_slots(const dsm_sched* h);
int64_t dsm_sched_step_idx(const dsm_sched* h);
/* The mask [B] and the gathered batch [B,1920] of the last tick (host memory; for tests). */
const uint8_t* dsm_sched_last_mask(const dsm_sched* h);
const float* dsm_sched_batch_pcm(const dsm_sched* h);

/* ---- weight and configuration files (host side; moshi-server/src/batched_asr.rs:738-771) -------------------- */
/* A safetensors file mapped read-only: what candle's VarBuilder::from_mmaped_safetensors hands the reference's
 * constructors.  The table points INTO the mapping: keep the file open until the create call that consumes it has
 * returned (create copies to the device).  F32 / BF16 / F16 tensors; the load-time folds (weight-norm conv.rs:27-45,
 * codebooks quantization.rs:86-107) happen inside dsm_mimi_create. */
typedef struct dsm_weight_file dsm_weight_file;
dsm_status dsm_weights_open(const char* safetensors_path, dsm_weight_file** out);
const dsm_weights* dsm_weight_file_table(const dsm_weight_file* f);
/* __metadata__ entry `key` copied into buf; returns its length or -1 when absent. */
int64_t dsm_weight_file_metadata(const dsm_weight_file* f, const char* key, char* buf, size_t cap);
void dsm_weight_file_close(dsm_weight_file* f);
/* Writes a table as a safetensors file (tests, checkpoint conversion tooling). */
dsm_status dsm_weights_save(const char* safetensors_path, const dsm_weights* weights);

/* The reference's server TOML (configs/stt/config-stt-*.toml): [modules.<module>] + .model + .model.transformer +
 * .model.extra_heads -> the LM spec and the AsrConfig fields beside it.  Keys this library does not implement with
 * another value than the STT configs use (gating, norm, positional_embedding, kv_repeat, ...) are refused. */
typedef struct dsm_asr_config {
  dsm_lm_spec lm;
  int32_t asr_delay_in_tokens;      /* or conditioning_delay * 12.5 + 1 when that key is present (batched_asr.rs:769-770) */
  int32_t has_conditioning_delay;
  float conditioning_delay;
  float temperature;
  int32_t batch_size;
  char module_type[32];             /* "BatchedAsr" / "Asr" */
  char lm_model_file[512];
  char audio_tokenizer_file[512];
  char text_tokenizer_file[512];
} dsm_asr_config;
dsm_status dsm_asr_config_from_toml(const char* toml_path, const char* module /* NULL = "asr" */, dsm_asr_config* out);
dsm_status dsm_asr_config_from_toml_text(const char* toml_text, const char* module, dsm_asr_config* out);

/* ---- wire messages (moshi-server/src/asr.rs:15-34) ------------------------------------------------------------ */
/* MessagePack bytes exactly as rmp_serde writes InMsg / OutMsg with `.with_struct_map()`: a map whose first entry is
 * "type" -> variant name, then the fields in declaration order.  Known-answer bytes:
 * client/rust/kyutai-client/src/stt/protocol.rs:81-96. */
enum { DSM_WIRE_IN_INIT = 0, DSM_WIRE_IN_MARKER = 1, DSM_WIRE_IN_AUDIO = 2, DSM_WIRE_IN_OGG_OPUS = 3, DSM_WIRE_IN_PING = 4 };
typedef struct dsm_wire_in {
  int32_t kind;
  int64_t id;           /* Marker */
  int64_t n;            /* Audio: samples; OggOpus: bytes */
  const float* pcm;     /* Audio */
  const uint8_t* data;  /* OggOpus (the Ogg/Opus decoder itself is outside this library) */
} dsm_wire_in;
enum { DSM_WIRE_OUT_WORD = 0, DSM_WIRE_OUT_END_WORD = 1, DSM_WIRE_OUT_MARKER = 2, DSM_WIRE_OUT_STEP = 3,
       DSM_WIRE_OUT_ERROR = 4, DSM_WIRE_OUT_READY = 5 };
typedef struct dsm_wire_out {
  int32_t kind;
  const char* text;     /* Word: text, Error: message (UTF-8, not NUL-terminated when decoded) */
  int32_t text_len;
  double time;          /* Word: start_time, EndWord: stop_time */
  int64_t id;           /* Marker */
  int64_t step_idx;     /* Step */
  const float* prs;     /* Step */
  int32_t n_prs;
  int64_t buffered_pcm; /* Step */
} dsm_wire_out;
/* *len receives the encoded size; DSM_ERR_CAPACITY when it exceeds cap. */
dsm_status dsm_wire_encode_out(const dsm_wire_out* msg, uint8_t* buf, size_t cap, size_t* len);
dsm_status dsm_wire_decode_out(const uint8_t* bytes, size_t len, dsm_wire_out* msg, float* prs_buf, int32_t prs_cap);
dsm_status dsm_wire_encode_in(const dsm_wire_in* msg, uint8_t* buf, size_t cap, size_t* len);
/* Audio samples are written to pcm_buf, OggOpus bytes to data_buf; msg->pcm / msg->data point at them. */
dsm_status dsm_wire_decode_in(const uint8_t* bytes, size_t len, dsm_wire_in* msg, float* pcm_buf, int64_t pcm_cap,
                              uint8_t* data_buf, int64_t data_cap);

/* ---- SentencePiece detokenisation (moshi-server/src/batched_asr.rs:667 decode_piece_ids) ---------------------- */
typedef struct dsm_spm dsm_spm;
dsm_status dsm_spm_open(const char* model_path, dsm_spm** out);
void dsm_spm_close(dsm_spm* h);
int32_t dsm_spm_vocab_size(const dsm_spm* h);
const char* dsm_spm_piece(const dsm_spm* h, int32_t id);
/* Token ids of one word -> UTF-8 text (NUL-terminated in buf); *len = length in bytes. */
dsm_status dsm_spm_decode(dsm_spm* h, const uint32_t* ids, int32_t n, char* buf, size_t cap, size_t* len);

/* ---- a connection's wire traffic through the scheduler (batched_asr.rs:894-985) ------------------------------- */
/* InMsg bytes in: Init re-arms reset + Ready, Audio / Marker / Ping as the push_* calls; OggOpus is DSM_ERR_UNSUPPORTED
 * (decode to PCM first: the Ogg/Opus decoder is outside this library). */
dsm_status dsm_sched_push_msg(dsm_sched* h, int32_t slot, uint64_t channel_id, const uint8_t* bytes, size_t len);
/* The connection's next queued event (threaded mode) as OutMsg bytes; Word text through `spm` (NULL: token ids). */
dsm_status dsm_sched_poll_msg(dsm_sched* h, int32_t slot, dsm_spm* spm, uint8_t* buf, size_t cap, size_t* len, int32_t* got);

/* ---- single-connection session: the non-batched Asr module (moshi-server/src/asr.rs:95-330) ------------------ */
/* One asr::State of batch size 1 driven by a connection's wire messages: Audio of any length is cut into frames and
 * stepped (Mimi encode + LM), AsrMsgs leave as OutMsg bytes with SentencePiece text (`spm` may be NULL: token ids),
 * and every step's tokens are kept for the log the reference writes at the end of a connection (asr.rs:126-176). */
typedef struct dsm_session dsm_session;
dsm_status dsm_session_create(dsm_asr* asr_batch1, dsm_spm* spm, dsm_session** out);
void dsm_session_destroy(dsm_session* s);
dsm_status dsm_session_push_msg(dsm_session* s, const uint8_t* bytes, size_t len);
dsm_status dsm_session_poll_msg(dsm_session* s, uint8_t* buf, size_t cap, size_t* len, int32_t* got);
int64_t dsm_session_steps(const dsm_session* s);
/* safetensors with "text" [1, steps] and "audio" [1, codebooks, steps], I64 */
dsm_status dsm_session_save_log(const dsm_session* s, const char* safetensors_path);

/* ---- TTS audio sink: the consumer of Mimi::decode_step (moshi-server/src/tts.rs:512-560, 917-929) -------------- */
/* A batch-1 Mimi decoder fed one step's audio tokens at a time: frames are decoded once
 * step_idx >= text_audio_delay_in_tokens + acoustic_delay and leave in the connection's format -- 0: raw
 * little-endian f32 (Encoder::Pcm), 1: MessagePack OutMsg::Audio{pcm} (Encoder::PcmMessagePack), tts.rs:252-281.
 * The text-to-token generation loop (LM with depformer) is not part of this library. */
typedef struct dsm_tts_sink dsm_tts_sink;
dsm_status dsm_tts_sink_create(dsm_mimi* mimi_decoder_batch1, int32_t text_audio_delay_in_tokens, int32_t acoustic_delay,
                               int32_t format, dsm_tts_sink** out);
void dsm_tts_sink_destroy(dsm_tts_sink* s);
/* codes [codebooks] of this step or NULL (a step without audio tokens); *len = bytes to send (0: none). */
dsm_status dsm_tts_sink_push(dsm_tts_sink* s, const uint32_t* codes, int64_t step_idx, uint8_t* buf, size_t cap, size_t* len);
/* A finished token matrix codes [codebooks, steps] -> pcm [steps * 1920] from a fresh decoder state. */
dsm_status dsm_tts_sink_decode_all(dsm_tts_sink* s, const uint32_t* codes, int64_t steps, float* pcm);

#ifdef __cplusplus
}
#endif
#endif /* DSM_B200_H */
