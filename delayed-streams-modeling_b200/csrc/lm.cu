// LM handle + ASR state.  Reference: moshi-core/src/{lm.rs:797-1008, batched_transformer.rs, kv_cache.rs, asr.rs}.
//
// Data layout (device):
//   * weights: bf16 row-major [out, in] exactly as stored (Linear = x . W^T), norm scales f32;
//   * residual stream x: f32 [B, D] (the reference keeps bf16; f32 is strictly closer to its f32 CPU
//     path, which is the parity oracle); GEMM operands (normalised x, attention out, gated MLP
//     activations) are bf16 like the reference's tensors;
//   * KV rings: per layer K and V as [B, H, ctx, hd] bf16 -- the reference's ScatteredKvCache shape
//     (kv_cache.rs:88-94) -- one contiguous ctx*hd run per (stream, head) so the attention kernel
//     streams each ring exactly once with 16-byte loads.
#include "lm.cuh"

#include <cmath>

#include "mimi.cuh"

namespace dsm {

static std::vector<float> rope_inv_freq_lm(int head_dim, float theta) {
  std::vector<float> v(head_dim / 2);
  for (int i = 0; i < head_dim; i += 2) v[i / 2] = 1.0f / powf(theta, (float)i / (float)head_dim);
  return v;
}

static uint64_t name_seed(const std::string& s) {
  uint64_t h = 1469598103934665603ULL;
  for (char c : s) h = (h ^ (unsigned char)c) * 1099511628211ULL;
  return h;
}

LmImpl::LmImpl(const dsm_lm_spec& sp, const dsm_weights* w, int dev, int batch, const dsm_lm_options* opt)
    : spec(sp), B(batch), device(dev) {
  DSM_CHECK(batch > 0, DSM_ERR_INVALID_ARG, "batch_size must be positive");
  if (opt) kv_dtype = opt->kv_dtype;
  DSM_CHECK(kv_dtype == DSM_KV_BF16 || kv_dtype == DSM_KV_INT8 || kv_dtype == DSM_KV_INT8K, DSM_ERR_INVALID_ARG,
            "unknown kv_dtype %d", kv_dtype);
  D = sp.d_model, H = sp.num_heads, L = sp.num_layers, ctx = sp.context;
  DSM_CHECK(D > 0 && H > 0 && D % H == 0, DSM_ERR_INVALID_ARG, "d_model %d not divisible by num_heads %d", D, H);
  hd = D / H;
  DSM_CHECK(hd == 64 || hd == 128, DSM_ERR_UNSUPPORTED, "head_dim %d (64 or 128)", hd);
  DSM_CHECK(D % 256 == 0, DSM_ERR_UNSUPPORTED, "d_model %d must be a multiple of 256", D);
  hidden = sp.dim_feedforward == 4 * D ? 11 * D / 4 : 2 * sp.dim_feedforward / 3;   // batched_transformer.rs:153-157
  DSM_CHECK(hidden % 16 == 0, DSM_ERR_UNSUPPORTED, "gated hidden size %d must be a multiple of 16", hidden);
  DSM_CUDA(cudaSetDevice(device));
  DSM_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  WeightTable wt(w);
  const bool random = (w == nullptr);

  auto bf16_mat = [&](const std::string& name, size_t n, float std) -> __nv_bfloat16* {
    weight_bytes += n * 2;
    if (random) {
      __nv_bfloat16* p = arena.alloc<__nv_bfloat16>(n, false);
      fill_random_bf16(p, n, std, name_seed(name), stream);
      return p;
    }
    return upload_tensor_bf16(arena, wt.find(name), n, stream);
  };
  auto f32_vec = [&](const std::string& name, size_t n) -> float* {
    weight_bytes += n * 4;
    if (random) {
      float* p = arena.alloc<float>(n, false);
      fill_random_f32(p, n, 0.1f, 1.0f, name_seed(name), stream);
      return p;
    }
    return arena.upload(wt.f32(name, n));
  };

  d_mask = arena.alloc<uint8_t>(B);
  ring.positions = arena.alloc<uint32_t>(B);
  ring.indices = arena.alloc<uint32_t>(B);
  ring.cur_index = arena.alloc<uint32_t>(B);
  ring.cur_pos = arena.alloc<uint32_t>(B);
  ring.rope_pos = arena.alloc<uint32_t>(B);
  inv_freq = arena.upload(rope_inv_freq_lm(hd, sp.max_period));

  const int n_cb = sp.audio_codebooks;
  const float emb_std = 1.f / sqrtf(1.f + (float)n_cb), sd = 1.f / sqrtf((float)D);
  text_emb = bf16_mat("text_emb.weight", (size_t)sp.text_in_vocab_size * D, emb_std);
  // all audio tables in one allocation so the embedding kernel indexes [codebook][id][D]
  {
    const size_t per = (size_t)sp.audio_vocab_size * D;
    audio_emb = arena.alloc<__nv_bfloat16>(per * n_cb, false);
    weight_bytes += per * n_cb * 2;
    for (int i = 0; i < n_cb; ++i) {
      const std::string name = "emb." + std::to_string(i) + ".weight";
      if (random) fill_random_bf16(audio_emb + per * i, per, emb_std, name_seed(name), stream);
      else {
        Arena tmp;
        __nv_bfloat16* t = upload_tensor_bf16(tmp, wt.find(name), per, stream);
        DSM_CUDA(cudaMemcpyAsync(audio_emb + per * i, t, per * 2, cudaMemcpyDeviceToDevice, stream));
        DSM_CUDA(cudaStreamSynchronize(stream));
      }
    }
  }
  layers.resize(L);
  for (int l = 0; l < L; ++l) {
    const std::string p = "transformer.layers." + std::to_string(l);
    Layer& y = layers[l];
    y.in_proj = bf16_mat(p + ".self_attn.in_proj_weight", (size_t)3 * D * D, sd);
    y.out_proj = bf16_mat(p + ".self_attn.out_proj.weight", (size_t)D * D, 0.5f * sd);
    y.norm1 = f32_vec(p + ".norm1.alpha", D);
    y.norm2 = f32_vec(p + ".norm2.alpha", D);
    y.lin_in = bf16_mat(p + ".gating.linear_in.weight", (size_t)2 * hidden * D, sd);
    y.lin_out = bf16_mat(p + ".gating.linear_out.weight", (size_t)D * hidden, 0.5f / sqrtf((float)hidden));
    const size_t ring_elems = (size_t)B * H * ctx * hd;
    y.k_ring = y.v_ring = nullptr;
    y.q8 = KvRingQ8{};
    if (kv_dtype != DSM_KV_BF16) {
      const int ctxp = (ctx + 3) & ~3;
      y.q8.ctxp = ctxp;
      y.q8.k = arena.alloc<int8_t>(ring_elems);
      y.q8.k_scale = arena.alloc<float>((size_t)B * H * ctxp);
      kv_bytes += ring_elems + (size_t)B * H * ctxp * 4;
      if (kv_dtype == DSM_KV_INT8) {
        y.q8.v = arena.alloc<int8_t>((size_t)B * H * ctxp * hd);
        y.q8.v_scale = arena.alloc<float>((size_t)B * H * ctxp);
        kv_bytes += (size_t)B * H * ctxp * (hd + 4);
      } else {
        y.q8.v_bf16 = arena.alloc<__nv_bfloat16>(ring_elems);
        kv_bytes += ring_elems * 2;
      }
    } else {
      y.k_ring = arena.alloc<__nv_bfloat16>(ring_elems);
      y.v_ring = arena.alloc<__nv_bfloat16>(ring_elems);
      kv_bytes += 2 * ring_elems * 2;
    }
  }
  out_norm = f32_vec("out_norm.alpha", D);
  text_linear = bf16_mat("text_linear.weight", (size_t)sp.text_out_vocab_size * D, 2.f * sd);
  if (sp.extra_heads_num > 0) {
    const size_t per = (size_t)sp.extra_heads_dim * D;
    extra_w = arena.alloc<__nv_bfloat16>(per * sp.extra_heads_num, false);
    weight_bytes += per * sp.extra_heads_num * 2;
    for (int i = 0; i < sp.extra_heads_num; ++i) {
      const std::string name = "extra_heads." + std::to_string(i) + ".weight";
      if (random) fill_random_bf16(extra_w + per * i, per, sd, name_seed(name), stream);
      else {
        Arena tmp;
        __nv_bfloat16* t = upload_tensor_bf16(tmp, wt.find(name), per, stream);
        DSM_CUDA(cudaMemcpyAsync(extra_w + per * i, t, per * 2, cudaMemcpyDeviceToDevice, stream));
        DSM_CUDA(cudaStreamSynchronize(stream));
      }
    }
  }

  d_text_ids = arena.alloc<uint32_t>(B);
  d_audio_ids = arena.alloc<uint32_t>((size_t)B * n_cb);
  d_cond = arena.alloc<float>(D);
  x = arena.alloc<float>((size_t)B * D);
  xn = arena.alloc<__nv_bfloat16>((size_t)B * D);
  qkv = arena.alloc<float>((size_t)B * 3 * D);
  q = arena.alloc<__nv_bfloat16>((size_t)B * D);
  attn = arena.alloc<__nv_bfloat16>((size_t)B * D);
  hbuf = arena.alloc<float>((size_t)B * 2 * hidden);
  gbuf = arena.alloc<__nv_bfloat16>((size_t)B * hidden);
  ys_bf16 = arena.alloc<__nv_bfloat16>((size_t)B * D);
  ys_f32 = arena.alloc<float>((size_t)B * D);
  logits = arena.alloc<float>((size_t)B * sp.text_out_vocab_size);
  d_sampled = arena.alloc<uint32_t>(B);

  {
    const char* e = getenv("DSM_LM_GEMM");
    use_tc = !(e && std::string(e) == "simt");
  }
  tune_plans(148);
  DSM_CUDA(cudaStreamSynchronize(stream));
}

// (Re)builds the GEMM launch configurations and both step plans for `sm_budget` SMs: tile shapes and split-K are
// chosen to fill the SMs the LM branch may use -- all of them, or its partition inside an ASR state (runtime.cuh).
void LmImpl::tune_plans(int sm_budget) {
  const dsm_lm_spec& sp = spec;
  sms = sm_budget;
  if (use_tc) {
    for (int l = 0; l < L; ++l) {
      Layer& y = layers[l];
      tc_gemm_prepare(y.g_qkv, y.in_proj, xn, B, 3 * D, D, TC_STORE_F32, qkv, nullptr, sms);
      tc_gemm_prepare(y.g_out, y.out_proj, attn, B, D, D, TC_ACCUM_F32, x, nullptr, sms);
      tc_gemm_prepare(y.g_in, y.lin_in, xn, B, 2 * hidden, D, TC_GATE_BF16, nullptr, gbuf, sms);
      tc_gemm_prepare(y.g_down, y.lin_out, gbuf, B, D, hidden, TC_ACCUM_F32, x, nullptr, sms);
    }
    tc_gemm_prepare(g_text, text_linear, ys_bf16, B, sp.text_out_vocab_size, D, TC_STORE_F32, logits, nullptr, sms);
    // one buffer for the split-K partial planes of all of them: each GEMM's partials are consumed by the very next
    // kernel of the stream (RMSNorm / RoPE-scatter / argmax) before the following GEMM writes its own
    std::vector<TcGemm*> all{&g_text};
    for (Layer& y : layers)
      for (TcGemm* g : {&y.g_qkv, &y.g_out, &y.g_in, &y.g_down}) all.push_back(g);
    size_t bytes = 0;
    for (TcGemm* g : all) bytes = std::max(bytes, g->partial_bytes);
    if (bytes > partial_buf_bytes) {
      partial_buf = arena.alloc<float>(bytes / sizeof(float), false);
      partial_buf_bytes = bytes;
    }
    for (TcGemm* g : all) g->partials = partial_buf;
  }
  plan.clear();
  plan_cond.clear();
  build_plan(plan, false);
  build_plan(plan_cond, true);
}

void LmImpl::move_to_partition(cudaStream_t st, int sm_budget) {
  DSM_CUDA(cudaSetDevice(device));
  DSM_CUDA(cudaStreamSynchronize(stream));
  retired_streams.push_back(stream);     // callers may still hold the old handle (dsm_lm_stream): it stays valid
  stream = st;
  graph.reset();                         // captured launches carry the old grids
  graph_cond.reset();
  if (sm_budget != sms) tune_plans(sm_budget);
}

LmImpl::~LmImpl() {
  cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);
  graph.reset();
  graph_cond.reset();
  for (auto& e : prof_events) {
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  if (stream) cudaStreamDestroy(stream);
  for (cudaStream_t s : retired_streams) cudaStreamDestroy(s);
}

static PartialSums partial_sums(const TcGemm& g, bool use_tc) {
  PartialSums ps;
  if (use_tc && g.split_k > 1) ps.planes = g.partials, ps.count = g.split_k, ps.stride = (long long)g.M * g.N;
  return ps;
}

void LmImpl::build_plan(Plan& p, bool with_cond) {
  LmImpl* s = this;
  const float* cond = with_cond ? d_cond : nullptr;
  p.push_back([=](cudaStream_t st) {
    launch_ring_advance(s->ring, s->d_mask, s->B, 1, s->ctx, st);
    launch_embed_sum(s->d_text_ids, s->d_audio_ids, s->spec.audio_codebooks, s->text_emb, s->spec.text_in_vocab_size,
                     s->audio_emb, s->spec.audio_vocab_size, cond, s->B, s->D, s->x, st);
  });
  // Split-K partial planes a residual projection left behind: the next RMSNorm folds them into x (gemm_tc.cuh).
  PartialSums pending;
  for (int l = 0; l < L; ++l) {
    const Layer y = layers[l];
    const PartialSums ps_in = pending, ps_qkv = partial_sums(y.g_qkv, use_tc), ps_out = partial_sums(y.g_out, use_tc);
    pending = partial_sums(y.g_down, use_tc);
    p.push_back([=](cudaStream_t st) {
      const int B = s->B, D = s->D, Hd = s->hidden;
      launch_rmsnorm(s->x, ps_in, y.norm1, s->spec.rms_eps, B, D, s->xn, nullptr, st);
      if (s->use_tc) {
        tc_gemm_launch(y.g_qkv, st);
      } else {
        GemmParams g;
        g.A = s->xn, g.a_map = dense_rows(1, D), g.W = y.in_proj, g.M = B, g.N = 3 * D, g.K = D;
        g.out_raw = s->qkv, g.o_map = dense_rows(1, 3 * D);
        launch_gemm_simt<__nv_bfloat16, __nv_bfloat16>(g, st);
      }
      if (s->kv_dtype != DSM_KV_BF16) {
        launch_rope_scatter_q8(s->qkv, ps_qkv, s->ring, B, s->H, s->hd, s->ctx, s->inv_freq, s->q, y.q8, st);
        s->profile_begin(st);
        launch_lm_attention_q8(s->q, s->ring, s->d_mask, B, s->H, s->hd, s->ctx, y.q8, s->attn, st);
        s->profile_end(st);
      } else {
        launch_rope_scatter_bf16(s->qkv, ps_qkv, s->ring, B, s->H, s->hd, s->ctx, s->inv_freq, s->q, y.k_ring, y.v_ring, st);
        s->profile_begin(st);
        launch_lm_attention(s->q, s->ring, s->d_mask, B, s->H, s->hd, s->ctx, y.k_ring, y.v_ring, s->attn, st);
        s->profile_end(st);
      }
      if (s->use_tc) {
        tc_gemm_launch(y.g_out, st);       // x += out_proj(attn)   (or its partial planes)
      } else {
        GemmParams o;
        o.A = s->attn, o.a_map = dense_rows(1, D), o.W = y.out_proj, o.M = B, o.N = D, o.K = D;
        o.res = s->x, o.r_map = dense_rows(1, D), o.out_raw = s->x, o.o_map = dense_rows(1, D);
        launch_gemm_simt<__nv_bfloat16, __nv_bfloat16>(o, st);
      }
      launch_rmsnorm(s->x, ps_out, y.norm2, s->spec.rms_eps, B, D, s->xn, nullptr, st);
      if (s->use_tc) {
        tc_gemm_launch(y.g_in, st);        // gbuf = silu(gate) * value
        tc_gemm_launch(y.g_down, st);      // x += linear_out(gbuf)   (or its partial planes)
      } else {
        GemmParams f1;
        f1.A = s->xn, f1.a_map = dense_rows(1, D), f1.W = y.lin_in, f1.M = B, f1.N = 2 * Hd, f1.K = D;
        f1.out_raw = s->hbuf, f1.o_map = dense_rows(1, 2 * Hd);
        launch_gemm_simt<__nv_bfloat16, __nv_bfloat16>(f1, st);
        launch_silu_gate(s->hbuf, B, Hd, s->gbuf, st);
        GemmParams f2;
        f2.A = s->gbuf, f2.a_map = dense_rows(1, Hd), f2.W = y.lin_out, f2.M = B, f2.N = D, f2.K = Hd;
        f2.res = s->x, f2.r_map = dense_rows(1, D), f2.out_raw = s->x, f2.o_map = dense_rows(1, D);
        launch_gemm_simt<__nv_bfloat16, __nv_bfloat16>(f2, st);
      }
    });
  }
  const PartialSums ps_last = pending, ps_text = partial_sums(g_text, use_tc);
  p.push_back([=](cudaStream_t st) {
    const int B = s->B, D = s->D, V = s->spec.text_out_vocab_size;
    launch_rmsnorm(s->x, ps_last, s->out_norm, s->spec.rms_eps, B, D, s->ys_bf16, s->ys_f32, st);
    if (s->use_tc) {
      tc_gemm_launch(s->g_text, st);
    } else {
      GemmParams g;
      g.A = s->ys_bf16, g.a_map = dense_rows(1, D), g.W = s->text_linear, g.M = B, g.N = V, g.K = D;
      g.out_raw = s->logits, g.o_map = dense_rows(1, V);
      launch_gemm_simt<__nv_bfloat16, __nv_bfloat16>(g, st);
    }
    // greedy token (asr.rs:208-210); also where the text head's partial planes become the logits
    launch_argmax(s->logits, ps_text, B, V, s->d_sampled, st);
  });
}

static bool stream_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  return cs != cudaStreamCaptureStatusNone;
}
void LmImpl::profile_begin(cudaStream_t st) {
  if (!profile_attn || stream_capturing(st)) return;
  if (prof_used == prof_events.size()) {
    cudaEvent_t a, b;
    DSM_CUDA(cudaEventCreate(&a));
    DSM_CUDA(cudaEventCreate(&b));
    prof_events.emplace_back(a, b);
  }
  DSM_CUDA(cudaEventRecord(prof_events[prof_used].first, st));
}
void LmImpl::profile_end(cudaStream_t st) {
  if (!profile_attn || stream_capturing(st)) return;
  DSM_CUDA(cudaEventRecord(prof_events[prof_used].second, st));
  ++prof_used;
}
void LmImpl::profile_read(double* ms_sum, long long* launches) {
  DSM_CUDA(cudaSetDevice(device));
  DSM_CUDA(cudaStreamSynchronize(stream));
  double tot = 0.0;
  for (size_t i = 0; i < prof_used; ++i) {
    float ms = 0.f;
    DSM_CUDA(cudaEventElapsedTime(&ms, prof_events[i].first, prof_events[i].second));
    tot += ms;
  }
  if (ms_sum) *ms_sum = tot;
  if (launches) *launches = (long long)prof_used;
  prof_used = 0;
}

void LmImpl::upload_mask(const uint8_t* host_mask, cudaStream_t st) {
  if (host_mask) DSM_CUDA(cudaMemcpyAsync(d_mask, host_mask, B, cudaMemcpyHostToDevice, st));
  else DSM_CUDA(cudaMemsetAsync(d_mask, 1, B, st));
}

void LmImpl::forward(const uint32_t* text_ids, const uint32_t* audio_ids, const float* cond, const uint8_t* mask,
                     float* logits_out, float* ys_out) {
  DSM_CHECK(text_ids && audio_ids, DSM_ERR_INVALID_ARG, "text_ids and audio_ids are required");
  DSM_CUDA(cudaSetDevice(device));
  upload_mask(mask, stream);
  DSM_CUDA(cudaMemcpyAsync(d_text_ids, text_ids, (size_t)B * 4, cudaMemcpyDefault, stream));
  DSM_CUDA(cudaMemcpyAsync(d_audio_ids, audio_ids, (size_t)B * spec.audio_codebooks * 4, cudaMemcpyDefault, stream));
  if (cond) DSM_CUDA(cudaMemcpyAsync(d_cond, cond, (size_t)D * 4, cudaMemcpyDefault, stream));
  Plan& p = cond ? plan_cond : plan;
  GraphExec& g = cond ? graph_cond : graph;
  if (graphs_enabled()) {
    if (!g.ready()) g.capture(stream, [&] { run_plan(p, stream); });
    g.launch(stream);
  } else {
    run_plan(p, stream);
  }
  bool need_sync = false;
  if (logits_out) {
    DSM_CUDA(cudaMemcpyAsync(logits_out, logits, (size_t)B * spec.text_out_vocab_size * 4, cudaMemcpyDefault, stream));
    need_sync |= !is_device_pointer(logits_out);
  }
  if (ys_out) {
    DSM_CUDA(cudaMemcpyAsync(ys_out, ys_f32, (size_t)B * D * 4, cudaMemcpyDefault, stream));
    need_sync |= !is_device_pointer(ys_out);
  }
  if (need_sync) DSM_CUDA(cudaStreamSynchronize(stream));
}

void LmImpl::reset_state() {
  DSM_CUDA(cudaSetDevice(device));
  DSM_CUDA(cudaMemsetAsync(ring.positions, 0, B * 4, stream));
  DSM_CUDA(cudaMemsetAsync(ring.indices, 0, B * 4, stream));
  DSM_CUDA(cudaStreamSynchronize(stream));
}

void LmImpl::reset_batch_idx(int b) {
  DSM_CHECK(b >= 0 && b < B, DSM_ERR_INVALID_ARG, "batch_idx %d out of range [0, %d)", b, B);
  DSM_CUDA(cudaSetDevice(device));
  DSM_CUDA(cudaMemsetAsync(ring.positions + b, 0, 4, stream));
  DSM_CUDA(cudaMemsetAsync(ring.indices + b, 0, 4, stream));
}

// ================================================================================================
// ASR state (asr.rs:53-284)
// ================================================================================================
AsrImpl::AsrImpl(MimiImpl* mi, LmImpl* l, int dly, float temperature) : mimi(mi), lm(l), delay(dly) {
  DSM_CHECK(lm != nullptr, DSM_ERR_INVALID_ARG, "lm handle is required");
  DSM_CHECK(temperature <= 0.f, DSM_ERR_UNSUPPORTED, "only greedy sampling (temperature <= 0) is implemented");
  B = lm->B, n_cb = lm->spec.audio_codebooks, nh = lm->spec.extra_heads_num;
  if (mimi) {
    DSM_CHECK(mimi->B == B, DSM_ERR_SHAPE, "batch size mismatch: mimi %d != lm %d", mimi->B, B);
    DSM_CHECK(mimi->spec.n_q == n_cb, DSM_ERR_SHAPE, "mimi n_q %d != lm audio_codebooks %d", mimi->spec.n_q, n_cb);
    DSM_CHECK(mimi->device == lm->device, DSM_ERR_INVALID_ARG, "mimi and lm live on different devices");
  }
  DSM_CUDA(cudaSetDevice(lm->device));
  next_cb = arena.alloc<uint32_t>((size_t)B * n_cb);
  text_tok = arena.alloc<uint32_t>(B);
  first_step = arena.alloc<uint8_t>(B);
  codes_in = arena.alloc<uint32_t>((size_t)B * n_cb);
  sampled = lm->d_sampled;          // greedy tokens: the last kernel of the LM plan
  prs = arena.alloc<float>((size_t)std::max(nh, 1) * B);
  h_sampled = static_cast<uint32_t*>(arena.alloc_pinned((size_t)B * 4));
  h_prs = static_cast<float*>(arena.alloc_pinned((size_t)std::max(nh, 1) * B * 4));
  h_mask.assign(B, 1);
  items.resize(B);
  // SM partitions for the two branches of the frame graph (runtime.cuh): the handles move onto the partitions' streams
  // and re-tune their GEMM grids for the SMs they now own.
  if (mimi && mimi->spec.with_encoder) {
    const int want = sm_split_choice(B, lm->kv_dtype != DSM_KV_BF16);
    if (want > 0) {
      cudaStream_t s_lm = nullptr, s_codec = nullptr;
      SmPartitionInfo part;
      bool ok = true;
      try {
        part = create_partition_streams(lm->device, want, &s_lm, &s_codec);
      } catch (const Error& e) {
        // an optimisation, not a requirement (a driver without green contexts, a MIG slice): say so once and share the
        // SMs -- unless the size was asked for explicitly
        if (getenv("DSM_SM_SPLIT")) throw;
        fprintf(stderr, "dsm: SM partition unavailable (%s); the LM and codec branches share the device\n", e.what());
        ok = false;
      }
      if (ok) {
        lm->move_to_partition(s_lm, part.lm_sms);
        mimi->move_to_partition(s_codec, part.codec_sms);
        sm_split_lm = part.lm_sms, sm_split_codec = part.codec_sms;
      }
    }
  }
  for (cudaEvent_t* e : {&ev_fork, &ev_join}) DSM_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  for (cudaEvent_t* e : {&ev_t0, &ev_t1, &ev_m0, &ev_m1, &ev_l0, &ev_l1}) DSM_CUDA(cudaEventCreate(e));
  reset();
}

AsrImpl::~AsrImpl() {
  cudaSetDevice(lm->device);
  cudaStreamSynchronize(lm->stream);
  if (mimi) cudaStreamSynchronize(mimi->stream);
  graph_pcm.reset();
  graph_tokens.reset();
  for (cudaEvent_t e : {ev_fork, ev_join, ev_t0, ev_t1, ev_m0, ev_m1, ev_l0, ev_l1})
    if (e) cudaEventDestroy(e);
}

void AsrImpl::reset() {
  // asr.rs:104-113
  DSM_CUDA(cudaSetDevice(lm->device));
  lm->reset_state();
  if (mimi) mimi->reset_state();
  cudaStream_t st = lm->stream;
  fill_u32(next_cb, (size_t)B * n_cb, (uint32_t)(lm->spec.audio_vocab_size - 1), st);
  fill_u32(text_tok, B, (uint32_t)(lm->spec.text_in_vocab_size - 1), st);
  fill_u8(first_step, B, 1, st);
  DSM_CUDA(cudaStreamSynchronize(st));
  for (auto& it : items) it = ItemState{};
  model_step_idx = 0;
  msgs.clear();
}

void AsrImpl::reset_batch_idx(int b) {
  // asr.rs:257-266
  DSM_CHECK(b >= 0 && b < B, DSM_ERR_INVALID_ARG, "batch index out of range: %d >= %d", b, B);
  DSM_CUDA(cudaSetDevice(lm->device));
  items[b] = ItemState{};
  cudaStream_t st = lm->stream;
  fill_u8(first_step + b, 1, 1, st);
  fill_u32(text_tok + b, 1, (uint32_t)(lm->spec.text_in_vocab_size - 1), st);
  lm->reset_batch_idx(b);
  if (mimi) {
    mimi->reset_batch_idx(b);
    DSM_CUDA(cudaStreamSynchronize(mimi->stream));
  }
  DSM_CUDA(cudaStreamSynchronize(st));
}

// Enqueues one frame: the LM branch (on lm->stream) consumes the PREVIOUS frame's codes
// (asr.rs:172-176), so the Mimi encode of the current frame runs concurrently on mimi->stream and the
// two join only for the state commit.
void AsrImpl::enqueue(bool with_mimi, bool timing) {
  cudaStream_t sl = lm->stream;
  cudaStream_t sm = with_mimi ? mimi->stream : nullptr;
  const uint32_t pad = (uint32_t)(lm->spec.audio_vocab_size - 1), start = (uint32_t)(lm->spec.text_in_vocab_size - 1);
  if (with_mimi) {
    DSM_CUDA(cudaEventRecord(ev_fork, sl));
    DSM_CUDA(cudaStreamWaitEvent(sm, ev_fork, 0));
    if (timing) DSM_CUDA(cudaEventRecord(ev_m0, sm));
    run_plan(mimi->enc_plan, sm);
    if (timing) DSM_CUDA(cudaEventRecord(ev_m1, sm));
    DSM_CUDA(cudaEventRecord(ev_join, sm));
  }
  if (timing) DSM_CUDA(cudaEventRecord(ev_l0, sl));
  launch_asr_prepare(next_cb, first_step, text_tok, pad, start, B, n_cb, lm->d_audio_ids, lm->d_text_ids, sl);
  run_plan(lm->plan, sl);
  if (nh > 0)
    launch_extra_heads(lm->ys_bf16, lm->extra_w, nh, lm->spec.extra_heads_dim, B, lm->D, prs, sl);
  if (timing) DSM_CUDA(cudaEventRecord(ev_l1, sl));
  if (with_mimi) DSM_CUDA(cudaStreamWaitEvent(sl, ev_join, 0));
  const uint32_t* codes = with_mimi ? mimi->d_codes : codes_in;
  launch_asr_commit(codes, lm->d_mask, sampled, B, n_cb, next_cb, text_tok, first_step, sl);
}

void AsrImpl::step(const float* pcm, const uint32_t* codes, const uint8_t* mask, uint32_t* text_tokens) {
  step_begin(pcm, codes, mask);
  step_end(text_tokens);
}

// Enqueues one step without waiting for it: H2D of the inputs, the frame graph, D2H of the sampled tokens.  `pcm` /
// `codes` must stay untouched until step_end() (the batched scheduler double-buffers its pinned batch so that the
// gather of frame t+1 overlaps the device work of frame t).
void AsrImpl::step_begin(const float* pcm, const uint32_t* codes, const uint8_t* mask) {
  DSM_CHECK(!in_flight, DSM_ERR_INVALID_ARG, "step_begin: the previous step has not been finished with step_end");
  const bool with_mimi = pcm != nullptr;
  DSM_CHECK(with_mimi || codes, DSM_ERR_INVALID_ARG, "either pcm or codes is required");
  DSM_CHECK(!with_mimi || (mimi && mimi->spec.with_encoder), DSM_ERR_INVALID_ARG, "step_pcm needs a Mimi handle with an encoder");
  DSM_CUDA(cudaSetDevice(lm->device));
  cudaStream_t sl = lm->stream;
  for (int b = 0; b < B; ++b) h_mask[b] = mask ? (mask[b] ? 1 : 0) : 1;
  lm->upload_mask(mask ? h_mask.data() : nullptr, sl);
  if (with_mimi) {
    mimi->upload_mask(mask ? h_mask.data() : nullptr, sl);
    DSM_CUDA(cudaMemcpyAsync(mimi->d_pcm, pcm, (size_t)B * DSM_FRAME_SIZE * 4, cudaMemcpyDefault, sl));
  } else {
    DSM_CUDA(cudaMemcpyAsync(codes_in, codes, (size_t)B * n_cb * 4, cudaMemcpyDefault, sl));
  }
  DSM_CUDA(cudaEventRecord(ev_t0, sl));
  flight_graph = graphs_enabled();
  flight_mimi = with_mimi;
  if (flight_graph) {
    GraphExec& g = with_mimi ? graph_pcm : graph_tokens;
    if (!g.ready()) g.capture(sl, [&] { enqueue(with_mimi, false); });
    g.launch(sl);
  } else {
    enqueue(with_mimi, true);
  }
  DSM_CUDA(cudaEventRecord(ev_t1, sl));
  // the one device->host read of the step (asr.rs:202,217)
  DSM_CUDA(cudaMemcpyAsync(h_sampled, sampled, (size_t)B * 4, cudaMemcpyDeviceToHost, sl));
  if (nh > 0) DSM_CUDA(cudaMemcpyAsync(h_prs, prs, (size_t)nh * B * 4, cudaMemcpyDeviceToHost, sl));
  in_flight = true;
}

void AsrImpl::step_end(uint32_t* text_tokens) {
  DSM_CHECK(in_flight, DSM_ERR_INVALID_ARG, "step_end without step_begin");
  in_flight = false;
  DSM_CUDA(cudaSetDevice(lm->device));
  DSM_CUDA(cudaStreamSynchronize(lm->stream));
  DSM_CUDA(cudaEventElapsedTime(&last_ms[0], ev_t0, ev_t1));
  if (!flight_graph) {
    DSM_CUDA(cudaEventElapsedTime(&last_ms[2], ev_l0, ev_l1));
    if (flight_mimi) DSM_CUDA(cudaEventElapsedTime(&last_ms[1], ev_m0, ev_m1));
    else last_ms[1] = -1.f;
  } else {
    last_ms[1] = last_ms[2] = -1.f;
  }
  host_update(h_mask.data());
  if (text_tokens) memcpy(text_tokens, h_sampled, (size_t)B * 4);
}

// asr.rs:193-252: Step message with the VAD probabilities, then the per-slot word state machine.
void AsrImpl::host_update(const uint8_t* mask) {
  msgs.clear();
  model_step_idx += 1;
  if (nh > 0) {
    prs_copy.assign(h_prs, h_prs + (size_t)nh * B);
    msgs.push_back(AsrMsgStore{DSM_MSG_STEP, -1, model_step_idx, {}, 0.0});
  }
  for (int b = 0; b < B; ++b) {
    if (!mask[b]) continue;
    ItemState& it = items[b];
    const uint32_t tok = h_sampled[b];
    it.step_idx += 1;
    if (it.step_idx >= delay) {
      if (tok == 3 || tok == 0) {
        if (!it.word_tokens.empty()) {
          AsrMsgStore m{DSM_MSG_WORD, b, 0, {}, it.last_stop_time};
          m.tokens.swap(it.word_tokens);
          msgs.push_back(std::move(m));
          it.unended_word = true;
        }
      } else {
        it.word_tokens.push_back(tok);
      }
      if (tok == 0) {
        const double stop = (double)(it.step_idx - delay) / 12.5;
        if (it.unended_word) {
          it.unended_word = false;
          msgs.push_back(AsrMsgStore{DSM_MSG_END_WORD, b, 0, {}, stop});
        }
        it.last_stop_time = stop;
      }
    }
  }
}

}  // namespace dsm
