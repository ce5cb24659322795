// extern "C" boundary (include/dsm_b200.h): argument checks, exception -> status translation.
#include <cstring>

#include "handles.cuh"

using namespace dsm;

template <typename F>
static dsm_status guarded(F&& f) {
  try {
    f();
    return DSM_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::bad_alloc&) {
    set_last_error("host allocation failed");
    return DSM_ERR_OOM;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return DSM_ERR_INVALID_ARG;
  } catch (...) {
    set_last_error("unknown error");
    return DSM_ERR_INVALID_ARG;
  }
}

#define REQUIRE(cond, msg) DSM_CHECK(cond, DSM_ERR_INVALID_ARG, msg)

extern "C" {

const char* dsm_version(void) { return "dsm_b200 0.1.0 (sm_100a)"; }

size_t dsm_last_error(char* buf, size_t cap) {
  const std::string& s = last_error();
  if (buf && cap) {
    size_t n = std::min(cap - 1, s.size());
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return s.size();
}

int32_t dsm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void dsm_set_graphs(int32_t enabled) { set_graphs_enabled(enabled); }
long long dsm_launch_count(void) { return g_launch_count; }
void dsm_launch_count_reset(void) { g_launch_count = 0; }

void dsm_mimi_spec_v0_1(dsm_mimi_spec* s, int32_t n_q) {
  if (!s) return;
  memset(s, 0, sizeof *s);
  s->n_q = n_q > 0 ? n_q : 16;   // mimi.rs:88: default 16 when unspecified
  s->bins = 2048, s->qdim = 256, s->dim = 512, s->n_filters = 64;
  s->ratios[0] = 8, s->ratios[1] = 6, s->ratios[2] = 5, s->ratios[3] = 4;
  s->kernel_size = 7, s->residual_kernel_size = 3, s->last_kernel_size = 3, s->compress = 2;
  s->num_heads = 8, s->num_layers = 8, s->dim_feedforward = 2048, s->context = 250;
  s->max_period = 10000.f, s->ln_eps = 1e-5f, s->downsample_stride = 2;
  s->with_encoder = 1, s->with_decoder = 1;
}

void dsm_lm_spec_stt_1b_en_fr(dsm_lm_spec* s) {
  if (!s) return;
  memset(s, 0, sizeof *s);
  s->d_model = 2048, s->num_heads = 16, s->num_layers = 16, s->dim_feedforward = 8192, s->context = 750;
  s->max_period = 100000.f, s->text_in_vocab_size = 8001, s->text_out_vocab_size = 8000;
  s->audio_vocab_size = 2049, s->audio_codebooks = 32, s->extra_heads_num = 4, s->extra_heads_dim = 6;
  s->rms_eps = 1e-8f;
}

void dsm_lm_spec_stt_2_6b_en(dsm_lm_spec* s) {
  if (!s) return;
  memset(s, 0, sizeof *s);
  s->d_model = 2048, s->num_heads = 32, s->num_layers = 48, s->dim_feedforward = 8192, s->context = 375;
  s->max_period = 100000.f, s->text_in_vocab_size = 4001, s->text_out_vocab_size = 4000;
  s->audio_vocab_size = 2049, s->audio_codebooks = 32, s->extra_heads_num = 0, s->extra_heads_dim = 0;
  s->rms_eps = 1e-8f;
}

static void require_device(int32_t device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    fail(DSM_ERR_CUDA, "no CUDA device available (%s); dsm_b200 has no CPU fallback", cudaGetErrorString(e));
  }
  DSM_CHECK(device >= 0 && device < n, DSM_ERR_INVALID_ARG, "device %d out of range [0, %d)", device, n);
  cudaDeviceProp p;
  DSM_CUDA(cudaGetDeviceProperties(&p, device));
  DSM_CHECK(p.major == 10, DSM_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
            p.major, p.minor);
}

// Stand-alone GEMM check: out = X[M,K] . W[N,K]^T (bf16 bit patterns in, f32 accumulate) through the
// CUDA-core kernel (impl 0) or the tcgen05 kernel (impl 1).  mode: 0 store f32 [M,N]; 1 accumulate into
// inout_f32 [M,N]; 2 gated pair -> out_bf16 [M, N/2] = silu(acc[:, j]) * acc[:, N/2 + j].
dsm_status dsm_debug_gemm_bf16(int32_t impl, int32_t mode, int32_t M, int32_t N, int32_t K, const uint16_t* X,
                               const uint16_t* W, float* inout_f32, uint16_t* out_bf16) {
  return guarded([&] {
    REQUIRE(X && W, "X and W are required");
    require_device(0);
    DSM_CUDA(cudaSetDevice(0));
    Arena a;
    __nv_bfloat16* dX = a.alloc<__nv_bfloat16>((size_t)M * K);
    __nv_bfloat16* dW = a.alloc<__nv_bfloat16>((size_t)N * K);
    float* dO = a.alloc<float>((size_t)M * N);
    __nv_bfloat16* dG = a.alloc<__nv_bfloat16>((size_t)M * (N / 2 + 1));
    DSM_CUDA(cudaMemcpy(dX, X, (size_t)M * K * 2, cudaMemcpyHostToDevice));
    DSM_CUDA(cudaMemcpy(dW, W, (size_t)N * K * 2, cudaMemcpyHostToDevice));
    if (mode == 1) {
      REQUIRE(inout_f32, "inout_f32 is required");
      DSM_CUDA(cudaMemcpy(dO, inout_f32, (size_t)M * N * 4, cudaMemcpyHostToDevice));
    }
    if (impl == 1) {
      TcGemm g;
      tc_gemm_prepare(g, dW, dX, M, N, K, mode, dO, dG);
      if (g.partial_bytes) g.partials = a.alloc<float>(g.partial_bytes / sizeof(float), false);
      tc_gemm_launch(g, 0);
      tc_reduce_partials(g, 0);
    } else {
      GemmParams g;
      g.A = dX, g.a_map = dense_rows(1, K), g.W = dW, g.M = M, g.N = N, g.K = K;
      g.out_raw = dO, g.o_map = dense_rows(1, N);
      if (mode == 1) g.res = dO, g.r_map = dense_rows(1, N);
      launch_gemm_simt<__nv_bfloat16, __nv_bfloat16>(g, 0);
      if (mode == 2) launch_silu_gate(dO, M, N / 2, dG, 0);
    }
    DSM_CUDA(cudaDeviceSynchronize());
    if (mode == 2) {
      REQUIRE(out_bf16, "out_bf16 is required");
      DSM_CUDA(cudaMemcpy(out_bf16, dG, (size_t)M * (N / 2) * 2, cudaMemcpyDeviceToHost));
    } else {
      REQUIRE(inout_f32, "inout_f32 is required");
      DSM_CUDA(cudaMemcpy(inout_f32, dO, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
    }
  });
}

// Stand-alone check of the codec GEMM kernels: out[(b,t), n] = sum_k A[b*stride_b + t*stride_t + k] * W[n, k]
// (the implicit-GEMM conv window form), f32 in / f32 out, through the CUDA-core kernel (impl 0) or the
// tcgen05 split-bf16 kernel (impl 1).
dsm_status dsm_debug_gemm_f32(int32_t impl, int32_t B, int32_t T, int64_t stride_b, int64_t stride_t, int32_t N, int32_t K,
                              const float* A, int64_t a_elems, const float* W, const float* bias, float* out) {
  return guarded([&] {
    REQUIRE(A && W && out, "A, W and out are required");
    require_device(0);
    DSM_CUDA(cudaSetDevice(0));
    Arena a;
    float* dA = a.alloc<float>((size_t)a_elems);
    float* dO = a.alloc<float>((size_t)B * T * N);
    DSM_CUDA(cudaMemcpy(dA, A, (size_t)a_elems * 4, cudaMemcpyHostToDevice));
    std::vector<float> hw(W, W + (size_t)N * K);
    GemmParams g;
    g.A = dA, g.a_map = RowMap{T, (long long)stride_b, (long long)stride_t}, g.M = B * T, g.N = N, g.K = K;
    g.out_raw = dO, g.o_map = dense_rows(1, N);
    if (bias) g.bias = a.upload(std::vector<float>(bias, bias + N));
    __nv_bfloat16* dOP = nullptr;
    if (impl == 2 || impl == 3) {
      // A handed over as three bf16 planes (exact truncation split, as the producers of the codec do it)
      std::vector<uint16_t> pl((size_t)3 * a_elems);
      for (int64_t i = 0; i < a_elems; ++i) {
        float x = A[i];
        for (int q = 0; q < 3; ++q) {
          uint32_t u;
          memcpy(&u, &x, 4);
          u &= 0xffff0000u;
          float hi;
          memcpy(&hi, &u, 4);
          pl[(size_t)q * a_elems + i] = (uint16_t)(u >> 16);
          x -= hi;
        }
      }
      __nv_bfloat16* dP = a.alloc<__nv_bfloat16>(pl.size());
      DSM_CUDA(cudaMemcpy(dP, pl.data(), pl.size() * 2, cudaMemcpyHostToDevice));
      g.a_planes = dP, g.a_plane_stride = a_elems;
      if (impl == 3) {
        dOP = a.alloc<__nv_bfloat16>((size_t)3 * B * T * N);
        g.out_planes = dOP, g.p_map = dense_rows(1, N), g.out_plane_stride = (long long)B * T * N;
      }
    }
    if (impl >= 1) {
      DSM_CHECK(tc6_supported(g), DSM_ERR_UNSUPPORTED, "shape not supported by the tensor-core codec GEMM");
      Tc6Gemm t;
      tc6_prepare(t, g, tc6_split_weights(a, hw, N, K), tc6_make_workspace(a));
      tc6_launch(t, 0);
    } else {
      g.W = a.upload(hw);
      launch_gemm_simt<float, float>(g, 0);
    }
    DSM_CUDA(cudaDeviceSynchronize());
    DSM_CUDA(cudaMemcpy(out, dO, (size_t)B * T * N * 4, cudaMemcpyDeviceToHost));
    if (impl == 3) {
      // return the sum of the stored planes: must reproduce the f32 result bit for bit
      const size_t n = (size_t)B * T * N;
      std::vector<uint16_t> pl(3 * n);
      DSM_CUDA(cudaMemcpy(pl.data(), dOP, pl.size() * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < n; ++i) {
        float acc = 0.f;
        for (int q = 0; q < 3; ++q) {
          uint32_t u = (uint32_t)pl[q * n + i] << 16;
          float f;
          memcpy(&f, &u, 4);
          acc += f;
        }
        DSM_CHECK(memcmp(&acc, &out[i], 4) == 0 || (acc != acc && out[i] != out[i]), DSM_ERR_CUDA,
                  "plane sum differs from the f32 result at %zu", i);
      }
    }
  });
}

dsm_status dsm_debug_ring_trace(int32_t batch, int32_t context, int32_t n_steps, const int32_t* t, const uint8_t* masks,
                                uint32_t* indices_out, float* mask_out) {
  return guarded([&] {
    REQUIRE(batch > 0 && context > 0 && n_steps >= 0 && t && masks && indices_out && mask_out, "invalid arguments");
    require_device(0);
    DSM_CUDA(cudaSetDevice(0));
    Arena a;
    RingState rs;
    rs.positions = a.alloc<uint32_t>(batch), rs.indices = a.alloc<uint32_t>(batch);
    rs.cur_index = a.alloc<uint32_t>(batch), rs.cur_pos = a.alloc<uint32_t>(batch), rs.rope_pos = a.alloc<uint32_t>(batch);
    uint8_t* d_mask = a.alloc<uint8_t>(batch);
    int t_max = 1;
    for (int i = 0; i < n_steps; ++i) {
      DSM_CHECK(t[i] > 0 && t[i] < context, DSM_ERR_INVALID_ARG, "step %d: t=%d outside (0, context)", i, t[i]);
      t_max = std::max(t_max, t[i]);
    }
    uint32_t* d_idx = a.alloc<uint32_t>((size_t)batch * t_max);
    float* d_am = a.alloc<float>((size_t)batch * t_max * context);
    for (int i = 0; i < n_steps; ++i) {
      DSM_CUDA(cudaMemcpy(d_mask, masks + (size_t)i * batch, batch, cudaMemcpyHostToDevice));
      launch_ring_advance(rs, d_mask, batch, t[i], context, 0);
      launch_ring_debug(rs, d_mask, batch, t[i], context, d_idx, d_am, 0);
      DSM_CUDA(cudaDeviceSynchronize());
      DSM_CUDA(cudaMemcpy(indices_out, d_idx, (size_t)batch * t[i] * 4, cudaMemcpyDeviceToHost));
      DSM_CUDA(cudaMemcpy(mask_out, d_am, (size_t)batch * t[i] * context * 4, cudaMemcpyDeviceToHost));
      indices_out += (size_t)batch * t[i];
      mask_out += (size_t)batch * t[i] * context;
    }
  });
}

// ---- Mimi ----------------------------------------------------------------------------------------
dsm_status dsm_mimi_create(const dsm_mimi_spec* spec, const dsm_weights* weights, int32_t device, int32_t batch_size,
                           dsm_mimi** out) {
  return guarded([&] {
    REQUIRE(spec && out, "spec and out are required");
    *out = nullptr;
    require_device(device);
    std::unique_ptr<MimiImpl> impl(new MimiImpl(*spec, weights, device, batch_size));
    *out = new dsm_mimi{impl.release()};
  });
}
void dsm_mimi_destroy(dsm_mimi* h) {
  if (!h) return;
  delete h->impl;
  delete h;
}
dsm_status dsm_mimi_encode_step(dsm_mimi* h, const float* pcm, const uint8_t* mask, uint32_t* codes, int32_t* produced) {
  return guarded([&] {
    REQUIRE(h && pcm, "handle and pcm are required");
    if (produced) *produced = 0;
    h->impl->encode_step(pcm, mask, codes);
    if (produced) *produced = 1;
  });
}
dsm_status dsm_mimi_decode_step(dsm_mimi* h, const uint32_t* codes, const uint8_t* mask, float* pcm, int32_t* produced) {
  return guarded([&] {
    REQUIRE(h && codes, "handle and codes are required");
    if (produced) *produced = 0;
    h->impl->decode_step(codes, mask, pcm);
    if (produced) *produced = 1;
  });
}
dsm_status dsm_mimi_last_latent(dsm_mimi* h, float* latent) {
  return guarded([&] {
    REQUIRE(h && latent, "handle and latent are required");
    MimiImpl* m = h->impl;
    REQUIRE(m->latent, "handle has no encoder");
    DSM_CUDA(cudaSetDevice(m->device));
    DSM_CUDA(cudaMemcpyAsync(latent, m->latent, (size_t)m->B * m->spec.dim * 4, cudaMemcpyDefault, m->stream));
    DSM_CUDA(cudaStreamSynchronize(m->stream));
  });
}
dsm_status dsm_mimi_debug_tensor(dsm_mimi* h, const char* name, float* out, int64_t cap, int64_t* numel) {
  return guarded([&] {
    REQUIRE(h && name, "handle and name are required");
    MimiImpl* m = h->impl;
    const float* src = nullptr;
    int64_t n = 0;
    const std::string s(name);
    if (s == "seanet_out") src = m->dbg_seanet, n = (int64_t)m->B * 2 * m->spec.dim;
    else if (s == "transformer_out") src = m->dbg_xfmr, n = (int64_t)m->B * 2 * m->spec.dim;
    else if (s == "latent") src = m->latent, n = (int64_t)m->B * m->spec.dim;
    else if (s == "dec_emb") src = m->dec_emb, n = (int64_t)m->B * m->spec.dim;
    else fail(DSM_ERR_INVALID_ARG, "unknown debug tensor '%s'", name);
    REQUIRE(src, "that tensor does not exist on this handle");
    if (numel) *numel = n;
    if (out) {
      DSM_CHECK(cap >= n, DSM_ERR_SHAPE, "buffer of %lld floats is too small for %lld", (long long)cap, (long long)n);
      DSM_CUDA(cudaSetDevice(m->device));
      DSM_CUDA(cudaMemcpyAsync(out, src, (size_t)n * 4, cudaMemcpyDefault, m->stream));
      DSM_CUDA(cudaStreamSynchronize(m->stream));
    }
  });
}
dsm_status dsm_mimi_reset_state(dsm_mimi* h) {
  return guarded([&] {
    REQUIRE(h, "handle is required");
    h->impl->reset_state();
  });
}
dsm_status dsm_mimi_reset_batch_idx(dsm_mimi* h, int32_t b) {
  return guarded([&] {
    REQUIRE(h, "handle is required");
    h->impl->reset_batch_idx(b);
  });
}
dsm_status dsm_mimi_synchronize(dsm_mimi* h) {
  return guarded([&] {
    REQUIRE(h, "handle is required");
    DSM_CUDA(cudaSetDevice(h->impl->device));
    DSM_CUDA(cudaStreamSynchronize(h->impl->stream));
  });
}
int32_t dsm_mimi_batch_size(const dsm_mimi* h) { return h ? h->impl->B : 0; }
int32_t dsm_mimi_n_q(const dsm_mimi* h) { return h ? h->impl->spec.n_q : 0; }
int64_t dsm_mimi_device_bytes(const dsm_mimi* h) { return h ? (int64_t)h->impl->arena.total_bytes() : 0; }
int64_t dsm_mimi_kv_bytes(const dsm_mimi* h) { return h ? (int64_t)(h->impl->enc_tr.kv_bytes + h->impl->dec_tr.kv_bytes) : 0; }

// ---- LM ------------------------------------------------------------------------------------------
dsm_status dsm_lm_create(const dsm_lm_spec* spec, const dsm_weights* weights, int32_t device, int32_t batch_size,
                         dsm_lm** out) {
  return guarded([&] {
    REQUIRE(spec && out, "spec and out are required");
    *out = nullptr;
    require_device(device);
    std::unique_ptr<LmImpl> impl(new LmImpl(*spec, weights, device, batch_size));
    *out = new dsm_lm{impl.release()};
  });
}
dsm_status dsm_lm_create_ex(const dsm_lm_spec* spec, const dsm_weights* weights, int32_t device, int32_t batch_size,
                            const dsm_lm_options* options, dsm_lm** out) {
  return guarded([&] {
    REQUIRE(spec && out, "spec and out are required");
    *out = nullptr;
    require_device(device);
    std::unique_ptr<LmImpl> impl(new LmImpl(*spec, weights, device, batch_size, options));
    *out = new dsm_lm{impl.release()};
  });
}
int32_t dsm_lm_kv_dtype(const dsm_lm* h) { return h ? h->impl->kv_dtype : -1; }
void dsm_lm_destroy(dsm_lm* h) {
  if (!h) return;
  delete h->impl;
  delete h;
}
dsm_status dsm_lm_forward(dsm_lm* h, const uint32_t* text_ids, const uint32_t* audio_ids, const float* cond,
                          const uint8_t* mask, float* logits, float* ys) {
  return guarded([&] {
    REQUIRE(h, "handle is required");
    h->impl->forward(text_ids, audio_ids, cond, mask, logits, ys);
  });
}
dsm_status dsm_lm_reset_state(dsm_lm* h) {
  return guarded([&] {
    REQUIRE(h, "handle is required");
    h->impl->reset_state();
  });
}
dsm_status dsm_lm_reset_batch_idx(dsm_lm* h, int32_t b) {
  return guarded([&] {
    REQUIRE(h, "handle is required");
    h->impl->reset_batch_idx(b);
    DSM_CUDA(cudaStreamSynchronize(h->impl->stream));
  });
}
dsm_status dsm_lm_synchronize(dsm_lm* h) {
  return guarded([&] {
    REQUIRE(h, "handle is required");
    DSM_CUDA(cudaSetDevice(h->impl->device));
    DSM_CUDA(cudaStreamSynchronize(h->impl->stream));
  });
}
int32_t dsm_lm_text_start_token(const dsm_lm* h) { return h ? h->impl->spec.text_in_vocab_size - 1 : -1; }
int32_t dsm_lm_audio_pad_token(const dsm_lm* h) { return h ? h->impl->spec.audio_vocab_size - 1 : -1; }
int32_t dsm_lm_in_audio_codebooks(const dsm_lm* h) { return h ? h->impl->spec.audio_codebooks : 0; }
void* dsm_lm_stream(const dsm_lm* h) { return h ? (void*)h->impl->stream : nullptr; }
void* dsm_mimi_stream(const dsm_mimi* h) { return h ? (void*)h->impl->stream : nullptr; }
dsm_status dsm_lm_profile_attention(dsm_lm* h, int32_t enable) {
  return guarded([&] {
    REQUIRE(h, "handle is required");
    h->impl->profile_attn = enable != 0;
    h->impl->prof_used = 0;
  });
}
dsm_status dsm_lm_profile_read(dsm_lm* h, double* ms_sum, long long* launches) {
  return guarded([&] {
    REQUIRE(h, "handle is required");
    h->impl->profile_read(ms_sum, launches);
  });
}
dsm_status dsm_lm_debug_set_position(dsm_lm* h, uint32_t pos) {
  return guarded([&] {
    REQUIRE(h, "handle is required");
    LmImpl* l = h->impl;
    DSM_CUDA(cudaSetDevice(l->device));
    fill_u32(l->ring.positions, l->B, pos, l->stream);
    fill_u32(l->ring.indices, l->B, pos % (uint32_t)l->ctx, l->stream);
    DSM_CUDA(cudaStreamSynchronize(l->stream));
  });
}
int64_t dsm_lm_kv_bytes(const dsm_lm* h) { return h ? (int64_t)h->impl->kv_bytes : 0; }
int64_t dsm_lm_weight_bytes(const dsm_lm* h) { return h ? (int64_t)h->impl->weight_bytes : 0; }

// ---- ASR -----------------------------------------------------------------------------------------
dsm_status dsm_asr_create(dsm_mimi* mimi, dsm_lm* lm, int32_t delay, float temperature, dsm_asr** out) {
  return guarded([&] {
    REQUIRE(lm && out, "lm and out are required");
    *out = nullptr;
    std::unique_ptr<AsrImpl> impl(new AsrImpl(mimi ? mimi->impl : nullptr, lm->impl, delay, temperature));
    *out = new dsm_asr{impl.release(), mimi, lm};
  });
}
void dsm_asr_destroy(dsm_asr* h) {
  if (!h) return;
  delete h->impl;
  delete h;
}
dsm_status dsm_asr_step_pcm(dsm_asr* h, const float* pcm, const uint8_t* mask, uint32_t* text_tokens) {
  return guarded([&] {
    REQUIRE(h && pcm, "handle and pcm are required");
    h->impl->step(pcm, nullptr, mask, text_tokens);
  });
}
dsm_status dsm_asr_step_tokens(dsm_asr* h, const uint32_t* codes, const uint8_t* mask, uint32_t* text_tokens) {
  return guarded([&] {
    REQUIRE(h && codes, "handle and codes are required");
    h->impl->step(nullptr, codes, mask, text_tokens);
  });
}
int32_t dsm_asr_msg_count(const dsm_asr* h) { return h ? (int32_t)h->impl->msgs.size() : 0; }
dsm_status dsm_asr_msg_get(const dsm_asr* h, int32_t i, dsm_asr_msg* out) {
  return guarded([&] {
    REQUIRE(h && out, "handle and out are required");
    const AsrImpl* a = h->impl;
    DSM_CHECK(i >= 0 && i < (int32_t)a->msgs.size(), DSM_ERR_INVALID_ARG, "message index %d out of range", i);
    const AsrMsgStore& m = a->msgs[i];
    memset(out, 0, sizeof *out);
    out->kind = m.kind, out->batch_idx = m.batch_idx, out->step_idx = m.step_idx, out->time = m.time;
    out->n_tokens = (int32_t)m.tokens.size();
    out->tokens = m.tokens.empty() ? nullptr : m.tokens.data();
    out->prs = m.kind == DSM_MSG_STEP ? a->prs_copy.data() : nullptr;
  });
}
dsm_status dsm_asr_reset(dsm_asr* h) {
  return guarded([&] {
    REQUIRE(h, "handle is required");
    h->impl->reset();
  });
}
dsm_status dsm_asr_reset_batch_idx(dsm_asr* h, int32_t b) {
  return guarded([&] {
    REQUIRE(h, "handle is required");
    h->impl->reset_batch_idx(b);
  });
}
int32_t dsm_asr_batch_size(const dsm_asr* h) { return h ? h->impl->B : 0; }
int32_t dsm_asr_model_step_idx(const dsm_asr* h) { return h ? h->impl->model_step_idx : 0; }
int32_t dsm_asr_delay_in_tokens(const dsm_asr* h) { return h ? h->impl->delay : 0; }
dsm_status dsm_asr_last_logits(dsm_asr* h, float* logits) {
  return guarded([&] {
    REQUIRE(h && logits, "handle and logits are required");
    LmImpl* l = h->impl->lm;
    DSM_CUDA(cudaSetDevice(l->device));
    DSM_CUDA(cudaMemcpyAsync(logits, l->logits, (size_t)l->B * l->spec.text_out_vocab_size * 4, cudaMemcpyDefault, l->stream));
    DSM_CUDA(cudaStreamSynchronize(l->stream));
  });
}
dsm_status dsm_asr_last_step_ms(dsm_asr* h, float out[3]) {
  return guarded([&] {
    REQUIRE(h && out, "handle and out are required");
    for (int i = 0; i < 3; ++i) out[i] = h->impl->last_ms[i];
  });
}
dsm_status dsm_asr_sm_partition(const dsm_asr* h, int32_t out[2]) {
  return guarded([&] {
    REQUIRE(h && out, "handle and out are required");
    out[0] = h->impl->sm_split_lm, out[1] = h->impl->sm_split_codec;
  });
}

}  // extern "C"
