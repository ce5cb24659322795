// Mimi handle implementation.  Reference: moshi-core/src/{mimi.rs,seanet.rs,conv.rs,quantization.rs,
// batched_transformer.rs,kv_cache.rs}.
//
// Data layout (all f32, device):
//   * activations are channels-last [B, T, C] so a conv's im2col window is one contiguous run of
//     k*C_in floats; every conv layer owns an input buffer [B, S + T_in, C_in] whose first
//     S = k - stride rows are the carried left context (StreamableConv1d::state_prev_xs,
//     conv.rs:231) and whose remaining rows the producer layer writes directly (no concat);
//   * buffers hold POST-ELU values when the consumer applies ELU to its input (ELU(0) = 0 keeps the
//     zero left pad valid), and a separate raw copy exists only where a resnet skip needs it;
//   * conv-transpose with k = 2*stride is computed as a k'=2 conv producing `stride` output frames
//     per input frame; its carried state is the previous input row instead of the reference's
//     partial output sums (conv.rs:459-477) -- the same numbers, added in a different order;
//   * the "state commit" at the end of a step copies, for ACTIVE streams only, the tail of each
//     buffer to its head (conv.rs:347-367: the mask only gates the state update).
#include "mimi.cuh"

#include <cmath>

namespace dsm {

using GetFn = std::function<std::vector<float>(const std::string&, size_t, float, float)>;

static std::vector<float> rope_inv_freq(int head_dim, float theta) {
  // transformer.rs:385-392: 1 / theta^(i / dim) for i = 0, 2, 4, ... in f32
  std::vector<float> v(head_dim / 2);
  for (int i = 0; i < head_dim; i += 2) v[i / 2] = 1.0f / powf(theta, (float)i / (float)head_dim);
  return v;
}

// ---------------------------------------------------------------------------------------------
// codec GEMM dispatch
// ---------------------------------------------------------------------------------------------
bool codec_gemm_tc_enabled() {
  const char* e = getenv("DSM_CODEC_GEMM");
  return !(e && std::string(e) == "simt");
}

// Skinny codec GEMMs (transformer, RVQ scores) take their A operand pre-split into bf16 planes written by the
// producing kernel; DSM_CODEC_PRESPLIT=0 keeps the f32 A path (converter warps) for A/B runs.
bool codec_presplit_enabled() {
  static const bool on = [] {
    const char* e = getenv("DSM_CODEC_PRESPLIT");
    return codec_gemm_tc_enabled() && !(e && e[0] == '0');
  }();
  return on;
}

static thread_local CodecGemmRegistry* g_registry = nullptr;

std::shared_ptr<CodecGemm> make_codec_gemm(Arena& arena, const GemmParams& g_in, const std::vector<float>& w_host,
                                           const Tc6Workspace& wsp) {
  auto c = std::make_shared<CodecGemm>();
  GemmParams g = g_in;
  if (!(codec_gemm_tc_enabled() && tc6_supported(g))) g.a_planes = nullptr, g.out_planes = nullptr;
  c->g = g;
  DSM_CHECK(w_host.size() == (size_t)g.N * g.K, DSM_ERR_SHAPE, "codec gemm: weight has %zu elements, expected %d x %d",
            w_host.size(), g.N, g.K);
  c->use_tc = codec_gemm_tc_enabled() && tc6_supported(g);
  // (Tried: the K <= 64 layers -- pure epilogue on the tensor-core kernel -- on the CUDA-core GEMM: slower at every
  // batch size, codec step 2.82 / 2.91 vs 2.79 ms at 256 streams, 14.1 / 15.1 vs 13.9 ms at 2,000.)
  if (c->use_tc) {
    c->w_planes = tc6_split_weights(arena, w_host, g.N, g.K);
    c->wsp = wsp;
    tc6_prepare(c->t, g, c->w_planes, wsp);
  } else {
    c->g.W = arena.upload(w_host);
  }
  if (g_registry) g_registry->all.push_back(c);
  return c;
}

void codec_gemm_set_registry(CodecGemmRegistry* r) { g_registry = r; }

void codec_gemms_retune(CodecGemmRegistry& r, int sms) {
  for (auto& c : r.all)
    if (c->use_tc) tc6_prepare(c->t, c->g, c->w_planes, c->wsp, sms);
}

void run_codec_gemm(const CodecGemm& c, cudaStream_t st) {
  if (c.use_tc) tc6_launch(c.t, st);
  else launch_gemm_simt<float, float>(c.g, st);
}

// ---------------------------------------------------------------------------------------------
// codec transformer
// ---------------------------------------------------------------------------------------------
void CodecTransformer::init(Arena& arena, const dsm_mimi_spec& s, int batch, const std::string& prefix,
                            const GetFn& get) {
  B = batch, T = 2, D = s.dim, H = s.num_heads, hd = D / H, ff = s.dim_feedforward, ctx = s.context, L = s.num_layers;
  eps = s.ln_eps;
  arena_ = &arena;
  DSM_CHECK(hd == 64 && D == 512, DSM_ERR_UNSUPPORTED, "codec transformer: d_model %d / heads %d (need 512 / 64-wide)", D, H);
  ring.positions = arena.alloc<uint32_t>(B);
  ring.indices = arena.alloc<uint32_t>(B);
  ring.cur_index = arena.alloc<uint32_t>(B);
  ring.cur_pos = arena.alloc<uint32_t>(B);
  ring.rope_pos = arena.alloc<uint32_t>(B);
  inv_freq = arena.upload(rope_inv_freq(hd, s.max_period));
  const size_t M = (size_t)B * T;
  xn = arena.alloc<float>(M * D);
  qkv = arena.alloc<float>(M * 3 * D);
  attn = arena.alloc<float>(M * D);
  hbuf = arena.alloc<float>(M * ff);
  if (codec_presplit_enabled()) {
    xn_pl = arena.alloc<__nv_bfloat16>(3 * M * D);
    attn_pl = arena.alloc<__nv_bfloat16>(3 * M * D);
    h_pl = arena.alloc<__nv_bfloat16>(3 * M * ff);
  }
  const float sd = 1.f / sqrtf((float)D), sff = 1.f / sqrtf((float)ff);
  layers.resize(L);
  for (int l = 0; l < L; ++l) {
    const std::string p = prefix + ".transformer.layers." + std::to_string(l);
    Layer& y = layers[l];
    y.h_in_proj = get(p + ".self_attn.in_proj_weight", (size_t)3 * D * D, sd, 0.f);
    y.h_out_proj = get(p + ".self_attn.out_proj.weight", (size_t)D * D, sd, 0.f);
    // LayerNorm weight is stored as `alpha` or `weight` (batched_transformer.rs:206-212)
    y.ln1w = arena.upload(get(p + ".norm1.alpha|" + p + ".norm1.weight", D, 0.1f, 1.f));
    y.ln1b = arena.upload(get(p + ".norm1.bias", D, 0.05f, 0.f));
    y.ln2w = arena.upload(get(p + ".norm2.alpha|" + p + ".norm2.weight", D, 0.1f, 1.f));
    y.ln2b = arena.upload(get(p + ".norm2.bias", D, 0.05f, 0.f));
    y.h_lin1 = get(p + ".linear1.weight", (size_t)ff * D, sd, 0.f);
    y.h_lin2 = get(p + ".linear2.weight", (size_t)D * ff, sff, 0.f);
    y.ls1 = arena.upload(get(p + ".layer_scale_1.scale", D, 0.f, 0.01f));
    y.ls2 = arena.upload(get(p + ".layer_scale_2.scale", D, 0.f, 0.01f));
    const size_t ring_elems = (size_t)B * H * ctx * hd;
    y.k_ring = arena.alloc<float>(ring_elems);
    y.v_ring = arena.alloc<float>(ring_elems);
    kv_bytes += 2 * ring_elems * sizeof(float);
  }
}

void CodecTransformer::append_plan(Plan& plan, float* x, RowMap x_map, const uint8_t* d_mask) {
  const int M = B * T;
  CodecTransformer* self = this;
  plan.push_back([=](cudaStream_t st) { launch_ring_advance(self->ring, d_mask, self->B, self->T, self->ctx, st); });
  for (int l = 0; l < L; ++l) {
    Layer& yl = layers[l];
    GemmParams g;     // qkv = in_proj(norm1(x))
    g.A = xn, g.a_map = dense_rows(1, D), g.M = M, g.N = 3 * D, g.K = D;
    g.out_raw = qkv, g.o_map = dense_rows(1, 3 * D);
    const long long MD = (long long)M * D, MF = (long long)M * ff;
    if (xn_pl) g.a_planes = xn_pl, g.a_plane_stride = MD;
    auto g_qkv = make_codec_gemm(*arena_, g, yl.h_in_proj, wsp);
    GemmParams o;     // x <- x + layer_scale_1 * out_proj(attn)     (batched_transformer.rs:347-349)
    o.A = attn, o.a_map = dense_rows(1, D), o.M = M, o.N = D, o.K = D;
    o.scale = yl.ls1, o.res = x, o.r_map = x_map, o.out_raw = x, o.o_map = x_map;
    if (attn_pl) o.a_planes = attn_pl, o.a_plane_stride = MD;
    auto g_out = make_codec_gemm(*arena_, o, yl.h_out_proj, wsp);
    GemmParams f1;    // h = gelu(linear1(norm2(x)))
    f1.A = xn, f1.a_map = dense_rows(1, D), f1.M = M, f1.N = ff, f1.K = D;
    f1.act = 1, f1.out_raw = hbuf, f1.o_map = dense_rows(1, ff);
    if (xn_pl) f1.a_planes = xn_pl, f1.a_plane_stride = MD;
    if (h_pl) f1.out_planes = h_pl, f1.p_map = dense_rows(1, ff), f1.out_plane_stride = MF;
    auto g_f1 = make_codec_gemm(*arena_, f1, yl.h_lin1, wsp);
    GemmParams f2;    // x <- x + layer_scale_2 * linear2(h)   (:360-361)
    f2.A = hbuf, f2.a_map = dense_rows(1, ff), f2.M = M, f2.N = D, f2.K = ff;
    f2.scale = yl.ls2, f2.res = x, f2.r_map = x_map, f2.out_raw = x, f2.o_map = x_map;
    if (h_pl && g_f1->use_tc) f2.a_planes = h_pl, f2.a_plane_stride = MF;   // planes exist only if linear1 wrote them
    auto g_f2 = make_codec_gemm(*arena_, f2, yl.h_lin2, wsp);
    yl.h_in_proj = std::vector<float>();
    yl.h_out_proj = std::vector<float>();
    yl.h_lin1 = std::vector<float>();
    yl.h_lin2 = std::vector<float>();
    const Layer y = yl;
    plan.push_back([=](cudaStream_t st) {
      launch_layernorm(x, x_map, y.ln1w, y.ln1b, self->eps, M, self->D, self->xn, st, self->xn_pl, MD);
      run_codec_gemm(*g_qkv, st);
      launch_rope_scatter_f32(self->qkv, self->ring, d_mask, self->B, self->T, self->H, self->hd, self->ctx,
                              self->inv_freq, y.k_ring, y.v_ring, st);
      launch_codec_attention(self->qkv, self->ring, d_mask, self->B, self->T, self->H, self->hd, self->ctx, y.k_ring,
                             y.v_ring, self->attn, st, self->attn_pl, MD);
      run_codec_gemm(*g_out, st);
      launch_layernorm(x, x_map, y.ln2w, y.ln2b, self->eps, M, self->D, self->xn, st, self->xn_pl, MD);
      run_codec_gemm(*g_f1, st);
      run_codec_gemm(*g_f2, st);
    });
  }
}

void CodecTransformer::reset_all(cudaStream_t st) {
  DSM_CUDA(cudaMemsetAsync(ring.positions, 0, B * sizeof(uint32_t), st));
  DSM_CUDA(cudaMemsetAsync(ring.indices, 0, B * sizeof(uint32_t), st));
}
void CodecTransformer::reset_batch_idx(int b, cudaStream_t st) {
  // kv_cache.rs:111-117: only the bookkeeping is zeroed; stale ring contents are masked out
  DSM_CUDA(cudaMemsetAsync(ring.positions + b, 0, sizeof(uint32_t), st));
  DSM_CUDA(cudaMemsetAsync(ring.indices + b, 0, sizeof(uint32_t), st));
}

// ---------------------------------------------------------------------------------------------
// weight preparation
// ---------------------------------------------------------------------------------------------
// conv.rs:27-45: weight, or weight_v * weight_g / ||weight_v||_(1,2) per leading index.  [n0][n1][k].
static std::vector<float> folded_weight(const GetFn& get, const WeightTable& wt, const std::string& p, int n0, int n1,
                                        int k, float std) {
  const size_t n = (size_t)n0 * n1 * k;
  if (wt.has(p + ".weight_g") && !wt.has(p + ".weight")) {
    std::vector<float> g = get(p + ".weight_g", n0, 0.f, 1.f);
    std::vector<float> v = get(p + ".weight_v", n, std, 0.f);
    const size_t per = (size_t)n1 * k;
    for (int o = 0; o < n0; ++o) {
      // accumulate in double: the closest a sequential sum gets to the reference's (pairwise, f32) one
      double ss = 0.0;
      for (size_t i = 0; i < per; ++i) ss += (double)v[o * per + i] * (double)v[o * per + i];
      const float norm = sqrtf((float)ss);
      for (size_t i = 0; i < per; ++i) v[o * per + i] = v[o * per + i] * g[o] / norm;
    }
    return v;
  }
  return get(p + ".weight", n, std, 0.f);
}

// [cout][cin][k] -> [cout][k][cin]  (GEMM K index = tap * cin + ci)
static std::vector<float> conv_to_gemm(const std::vector<float>& w, int cout, int cin, int k) {
  std::vector<float> o(w.size());
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int j = 0; j < k; ++j) o[((size_t)co * k + j) * cin + ci] = w[((size_t)co * cin + ci) * k + j];
  return o;
}

// conv-transpose [cin][cout][k = 2s] -> k'=2 conv weight [s*cout][2*cin]:
// row n = r*cout + co; tap 0 multiplies the previous input frame (kernel index r + s), tap 1 the
// current one (kernel index r).
static std::vector<float> convtr_to_gemm(const std::vector<float>& w, int cin, int cout, int k, int s) {
  std::vector<float> o((size_t)s * cout * 2 * cin);
  for (int r = 0; r < s; ++r)
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci) {
        const size_t row = (size_t)r * cout + co;
        o[row * 2 * cin + ci] = w[((size_t)ci * cout + co) * k + r + s];
        o[row * 2 * cin + cin + ci] = w[((size_t)ci * cout + co) * k + r];
      }
  return o;
}

static std::vector<SeanetLayer> encoder_layers(const dsm_mimi_spec& s) {
  // seanet.rs:160-245; indices of `encoder.model.N`
  std::vector<SeanetLayer> v;
  int mult = 1, idx = 1;
  const int nf = s.n_filters;
  v.push_back({"encoder.model.0.conv.conv", 1, nf, s.kernel_size, 1, false, 0, false});
  for (int i = 3; i >= 0; --i) {
    const int ratio = s.ratios[i], c = mult * nf, h = c / s.compress;
    v.push_back({"encoder.model." + std::to_string(idx) + ".block.1.conv.conv", c, h, s.residual_kernel_size, 1, true, 1, false});
    v.push_back({"encoder.model." + std::to_string(idx) + ".block.3.conv.conv", h, c, 1, 1, true, 2, false});
    v.push_back({"encoder.model." + std::to_string(idx + 2) + ".conv.conv", c, 2 * c, 2 * ratio, ratio, true, 0, false});
    idx += 3, mult *= 2;
  }
  v.push_back({"encoder.model." + std::to_string(idx + 1) + ".conv.conv", mult * nf, s.dim, s.last_kernel_size, 1, true, 0, false});
  return v;
}

static std::vector<SeanetLayer> decoder_layers(const dsm_mimi_spec& s) {
  // seanet.rs:312-399; indices of `decoder.model.N`
  std::vector<SeanetLayer> v;
  const int nf = s.n_filters;
  int mult = 16, idx = 1;
  v.push_back({"decoder.model.0.conv.conv", s.dim, mult * nf, s.kernel_size, 1, false, 0, false});
  for (int i = 0; i < 4; ++i) {
    const int ratio = s.ratios[i], c = mult * nf, c2 = c / 2, h = c2 / s.compress;
    v.push_back({"decoder.model." + std::to_string(idx + 1) + ".convtr.convtr", c, c2, 2 * ratio, ratio, true, 0, true});
    v.push_back({"decoder.model." + std::to_string(idx + 2) + ".block.1.conv.conv", c2, h, s.residual_kernel_size, 1, true, 1, false});
    v.push_back({"decoder.model." + std::to_string(idx + 2) + ".block.3.conv.conv", h, c2, 1, 1, true, 2, false});
    idx += 3, mult /= 2;
  }
  v.push_back({"decoder.model." + std::to_string(idx + 1) + ".conv.conv", nf, 1, s.last_kernel_size, 1, true, 0, false});
  return v;
}

// Builds the launch plan of a SEANet stack.
//  first_in: the first layer's input buffer [B, S0 + T0, cin0] (nullptr when cin0 == 1: pcm + state0 path)
//  final_out/final_map: where the last layer's raw output rows go (nullptr when cout_last == 1: pcm path)
struct SeanetBuild {
  MimiImpl* m;
  Arena* arena;
  int B;
  const GetFn* get;
  const WeightTable* wt;
  Plan* plan;
  std::vector<CommitEntry>* commits;
};

static void build_seanet(const SeanetBuild& sb, const std::vector<SeanetLayer>& L, int T0, float* first_in,
                         float* final_out, RowMap final_map, float* pcm_in, float* state0, float* pcm_out) {
  const int B = sb.B;
  const int n = (int)L.size();
  Arena& arena = *sb.arena;
  // geometry
  std::vector<int> T(n + 1), S(n), keff(n), seff(n), N(n);
  T[0] = T0;
  for (int i = 0; i < n; ++i) {
    if (L[i].tr) {
      DSM_CHECK(L[i].k == 2 * L[i].stride, DSM_ERR_UNSUPPORTED, "conv-transpose k=%d stride=%d (need k = 2*stride)", L[i].k, L[i].stride);
      keff[i] = 2, seff[i] = 1, S[i] = 1, N[i] = L[i].stride * L[i].cout, T[i + 1] = T[i] * L[i].stride;
    } else {
      DSM_CHECK(T[i] % L[i].stride == 0, DSM_ERR_SHAPE, "frame length %d not divisible by stride %d", T[i], L[i].stride);
      keff[i] = L[i].k, seff[i] = L[i].stride, S[i] = L[i].k - L[i].stride, N[i] = L[i].cout, T[i + 1] = T[i] / L[i].stride;
    }
  }
  // Encoder stage 1 (init conv + first resnet block) as one CUDA-core kernel (seanet_fused.cu) instead of an init
  // conv and two short-K GEMMs: codec step 2.44 vs 2.49 ms at 256 streams, 13.0-13.2 vs 13.7 ms at 2,000, equal at 64, and
  // 0.74 MB per stream less scratch.  DSM_SEANET_FUSED1=0 keeps the layer-by-layer path.
  static const bool fuse1_on = [] { const char* e = getenv("DSM_SEANET_FUSED1"); return !e || e[0] != '0'; }();
  const bool fuse1 = fuse1_on && !first_in && n > 3 && L[1].res == 1 && L[2].res == 2 && L[1].elu_in && L[2].elu_in && L[3].elu_in &&
                     !L[1].tr && !L[2].tr && L[1].cin == L[0].cout && L[2].cout == L[0].cout &&
                     seanet_stage1_supported(L[0].cout, L[0].k, L[1].cout, L[1].k, L[2].k, T0);
  float *f1_w0 = nullptr, *f1_w1t = nullptr, *f1_b1 = nullptr, *f1_w2t = nullptr;
  // input buffers
  std::vector<float*> X(n, nullptr);
  for (int i = 0; i < n; ++i) {
    if (i == 0) X[i] = first_in;
    else if (fuse1 && i == 2) continue;            // the block's hidden activation never leaves shared memory
    else X[i] = arena.alloc<float>((size_t)B * (S[i] + T[i]) * L[i].cin);
    if (X[i] && S[i] > 0) {
      const long long sb_ = (long long)(S[i] + T[i]) * L[i].cin;
      DSM_CHECK(S[i] * L[i].cin <= 4096, DSM_ERR_UNSUPPORTED, "conv state of %d floats exceeds the commit kernel's 4096", S[i] * L[i].cin);
      sb.commits->push_back(CommitEntry{X[i], X[i] + (long long)T[i] * L[i].cin, sb_, sb_, S[i] * L[i].cin, 0});
    }
  }
  if (!first_in) {
    // C_in == 1 first conv: left context lives in state0 [B, S0]; new samples in pcm_in [B, T0]
    sb.commits->push_back(CommitEntry{state0, pcm_in + (T0 - S[0]), S[0], T0, S[0], 0});
  }
  std::vector<float*> R(n, nullptr);   // raw copy of layer i's output when a resnet block starts at i+1
  int open_src = -1;                   // layer whose raw output the pending block adds back
  for (int i = 0; i < n; ++i) {
    const SeanetLayer& y = L[i];
    const bool last = i == n - 1;
    const float wstd = 1.4f / sqrtf((float)(y.cin * y.k)) * (y.tr ? sqrtf((float)y.stride) : 1.f);
    // ---- weights ----
    std::vector<float> hW;     // GEMM layout [N][K]
    float* dBias = nullptr;
    {
      std::vector<float> bias = (*sb.get)(y.wname + ".bias", y.cout, 0.05f, 0.f);
      if (y.tr) {
        std::vector<float> w = folded_weight(*sb.get, *sb.wt, y.wname, y.cin, y.cout, y.k, wstd);
        hW = convtr_to_gemm(w, y.cin, y.cout, y.k, y.stride);
        std::vector<float> b2((size_t)N[i]);
        for (int r = 0; r < y.stride; ++r)
          for (int co = 0; co < y.cout; ++co) b2[(size_t)r * y.cout + co] = bias[co];
        dBias = arena.upload(b2);
      } else {
        std::vector<float> w = folded_weight(*sb.get, *sb.wt, y.wname, y.cout, y.cin, y.k, wstd);
        hW = conv_to_gemm(w, y.cout, y.cin, y.k);
        dBias = arena.upload(bias);
      }
    }
    if (fuse1 && i < 2) {
      // weights in the fused kernel's layouts: w0 [64][7] as is, w1 [32][192] -> [192][32]
      if (i == 0) {
        std::vector<float> bias0 = (*sb.get)(y.wname + ".bias", y.cout, 0.05f, 0.f);
        std::vector<float> t((size_t)y.cout * 8, 0.f);             // [64][8]: the 7 taps, then the bias
        for (int co = 0; co < y.cout; ++co) {
          for (int q = 0; q < y.k; ++q) t[(size_t)co * 8 + q] = hW[(size_t)co * y.k + q];
          t[(size_t)co * 8 + 7] = bias0[co];
        }
        f1_w0 = arena.upload(t);
      }
      else {
        const int K = keff[i] * y.cin;
        std::vector<float> t((size_t)K * y.cout);
        for (int nn = 0; nn < y.cout; ++nn)
          for (int kk = 0; kk < K; ++kk) t[(size_t)kk * y.cout + nn] = hW[(size_t)nn * K + kk];
        f1_w1t = arena.upload(t), f1_b1 = dBias;
      }
      continue;
    }
    // ---- outputs ----
    float* out_act = nullptr;
    float* out_raw = nullptr;
    RowMap act_map{1, 0, 0}, raw_map{1, 0, 0};
    const int rows = y.tr ? T[i] : T[i + 1];     // GEMM rows per stream
    if (!last) {
      const SeanetLayer& nx = L[i + 1];
      const long long nb = (long long)(S[i + 1] + T[i + 1]) * nx.cin;
      float* dst = X[i + 1] + (long long)S[i + 1] * nx.cin;
      if (nx.elu_in) out_act = dst, act_map = RowMap{rows, nb, (long long)N[i]};
      else out_raw = dst, raw_map = RowMap{rows, nb, (long long)N[i]};
      if (nx.res == 1) {
        R[i] = arena.alloc<float>((size_t)B * T[i + 1] * y.cout);
        DSM_CHECK(out_raw == nullptr, DSM_ERR_UNSUPPORTED, "resnet block after a non-activated layer");
        out_raw = R[i], raw_map = RowMap{rows, (long long)T[i + 1] * y.cout, (long long)N[i]};
        open_src = i;
      }
    } else if (final_out) {
      out_raw = final_out, raw_map = final_map;
    }
    if (fuse1 && i == 2) {
      std::vector<float> t((size_t)y.cin * y.cout);               // w2 [64][32] -> [32][64]
      for (int nn = 0; nn < y.cout; ++nn)
        for (int kk = 0; kk < y.cin; ++kk) t[(size_t)kk * y.cout + nn] = hW[(size_t)nn * y.cin + kk];
      f1_w2t = arena.upload(t);
      DSM_CHECK(out_act && !out_raw, DSM_ERR_UNSUPPORTED, "fused stage 1 expects an activated consumer");
      float* x1 = X[1];
      const long long x1_stride = (long long)(S[1] + T[1]) * L[1].cin;
      const int T_in = T[0];
      sb.plan->push_back([=](cudaStream_t st) {
        launch_seanet_stage1(pcm_in, state0, f1_w0, f1_w1t, f1_b1, f1_w2t, dBias, x1, x1_stride, out_act, act_map, B, T_in, st);
      });
      continue;
    }
    const float* res = nullptr;
    RowMap res_map{1, 0, 0};
    if (y.res == 2) {
      DSM_CHECK(open_src >= 0 && R[open_src], DSM_ERR_UNSUPPORTED, "resnet close without open");
      res = R[open_src], res_map = RowMap{rows, (long long)T[i + 1] * y.cout, (long long)y.cout};
    }
    // ---- launch ----
    if (i == 0 && !first_in) {
      DSM_CHECK(y.cin == 1 && !y.tr, DSM_ERR_UNSUPPORTED, "first layer without an input buffer must have C_in = 1");
      const int S0 = S[0], T_in = T[0], C = y.cout, k = y.k;
      float* dW = arena.upload(hW);
      sb.plan->push_back([=](cudaStream_t st) {
        launch_init_conv(pcm_in, state0, S0, dW, dBias, B, T_in, C, k, out_raw, raw_map, out_act, act_map, st);
      });
    } else if (last && y.cout == 1 && !final_out) {
      DSM_CHECK(!y.tr, DSM_ERR_UNSUPPORTED, "C_out = 1 conv-transpose");
      const RowMap xm{T[i + 1], (long long)(S[i] + T[i]) * y.cin, (long long)seff[i] * y.cin};
      const float* xin = X[i];
      const int Tn = T[i + 1], C = y.cin, k = y.k;
      float* dW = arena.upload(hW);
      sb.plan->push_back([=](cudaStream_t st) { launch_final_conv(xin, xm, dW, dBias, B, Tn, C, k, pcm_out, st); });
    } else {
      GemmParams g;
      g.A = X[i];
      g.a_map = RowMap{rows, (long long)(S[i] + T[i]) * y.cin, (long long)seff[i] * y.cin};
      g.M = B * rows, g.N = N[i], g.K = keff[i] * y.cin;
      g.bias = dBias;
      g.res = res, g.r_map = res_map;
      g.out_raw = out_raw, g.o_map = raw_map;
      g.out_act = out_act, g.e_map = act_map;
      auto cg = make_codec_gemm(arena, g, hW, sb.m->wsp);
      sb.plan->push_back([=](cudaStream_t st) { run_codec_gemm(*cg, st); });
    }
  }
}

// ---------------------------------------------------------------------------------------------
// MimiImpl
// ---------------------------------------------------------------------------------------------
static uint64_t name_hash(const std::string& s) {
  uint64_t h = 1469598103934665603ULL;
  for (char c : s) h = (h ^ (unsigned char)c) * 1099511628211ULL;
  return h;
}

MimiImpl::MimiImpl(const dsm_mimi_spec& sp, const dsm_weights* w, int dev, int batch) : spec(sp), B(batch), device(dev) {
  DSM_CHECK(batch > 0, DSM_ERR_INVALID_ARG, "batch_size must be positive");
  DSM_CHECK(sp.n_q >= 1 && sp.n_q <= 32 && sp.qdim % 16 == 0 && sp.dim == 512, DSM_ERR_UNSUPPORTED,
            "mimi spec outside the v0_1 family (n_q=%d qdim=%d dim=%d)", sp.n_q, sp.qdim, sp.dim);
  DSM_CHECK(sp.with_encoder || sp.with_decoder, DSM_ERR_INVALID_ARG, "mimi spec enables neither encoder nor decoder");
  DSM_CUDA(cudaSetDevice(device));
  DSM_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  WeightTable wt(w);
  const bool random = (w == nullptr);
  // Tensor getter: from the table (either of `a|b` alternative names), or seeded host-side random init
  // of the same shape when no table was given (benchmarks).
  GetFn get = [&wt, random](const std::string& names, size_t n, float std, float mean) -> std::vector<float> {
    size_t bar = names.find('|');
    std::string a = bar == std::string::npos ? names : names.substr(0, bar);
    std::string b = bar == std::string::npos ? std::string() : names.substr(bar + 1);
    if (!random) {
      if (wt.has(a) || b.empty()) return wt.f32(a, n);
      return wt.f32(b, n);
    }
    std::vector<float> v(n);
    uint64_t s = name_hash(a) | 1;
    for (size_t i = 0; i < n; ++i) {
      s = s * 6364136223846793005ULL + 1442695040888963407ULL;
      const float u = (float)((s >> 40) & 0xffffff) * (1.f / 8388608.f) - 1.f;   // [-1, 1)
      v[i] = mean + std * 1.7320508f * u;
    }
    return v;
  };

  d_mask = arena.alloc<uint8_t>(B);
  h_mask = static_cast<uint8_t*>(arena.alloc_pinned(B));
  h_codes = static_cast<uint32_t*>(arena.alloc_pinned((size_t)B * spec.n_q * 4));
  d_first_step = arena.alloc<int>(1);
  d_first_step_dec = arena.alloc<int>(1);
  d_codes = arena.alloc<uint32_t>((size_t)B * spec.n_q);
  if (codec_gemm_tc_enabled()) wsp = tc6_make_workspace(arena);

  // ---- codebooks (shared by encode and decode), quantization.rs:86-107 ----
  {
    const int bins = spec.bins, qd = spec.qdim, nq = spec.n_q;
    std::vector<float> emb((size_t)nq * bins * qd), c2((size_t)nq * bins);
    for (int q = 0; q < nq; ++q) {
      const std::string p = q == 0 ? "quantizer.rvq_first.vq.layers.0._codebook"
                                   : "quantizer.rvq_rest.vq.layers." + std::to_string(q - 1) + "._codebook";
      std::vector<float> usage = get(p + ".cluster_usage", bins, 0.f, 1.f);
      std::vector<float> sum = get(p + ".embedding_sum", (size_t)bins * qd, 0.6f * powf(0.93f, (float)std::max(q - 1, 0)), 0.f);
      for (int j = 0; j < bins; ++j) {
        const float u = fmaxf(usage[j], 1e-5f);
        double acc = 0.0;
        for (int d = 0; d < qd; ++d) {
          const float e = sum[(size_t)j * qd + d] / u;
          emb[((size_t)q * bins + j) * qd + d] = e;
          acc += (double)(e * e);
        }
        c2[(size_t)q * bins + j] = (float)acc / 2.0f;
      }
    }
    emb_all = arena.upload(emb);
    c2_all = arena.upload(c2);
  }
  struct RegistryScope {
    explicit RegistryScope(CodecGemmRegistry* r) { codec_gemm_set_registry(r); }
    ~RegistryScope() { codec_gemm_set_registry(nullptr); }
  } scope(&gemms);
  if (spec.with_encoder) build_encoder(get, wt);
  if (spec.with_decoder) build_decoder(get, wt);
  reset_state();
  DSM_CUDA(cudaStreamSynchronize(stream));
}

MimiImpl::~MimiImpl() {
  cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);
  enc_graph.reset();
  dec_graph.reset();
  if (stream) cudaStreamDestroy(stream);
  for (cudaStream_t s : retired_streams) cudaStreamDestroy(s);
}

void MimiImpl::move_to_partition(cudaStream_t st, int sm_budget) {
  DSM_CUDA(cudaSetDevice(device));
  DSM_CUDA(cudaStreamSynchronize(stream));
  retired_streams.push_back(stream);     // callers may still hold the old handle (dsm_mimi_stream): it stays valid
  stream = st;
  enc_graph.reset();                     // captured launches carry the old grids
  dec_graph.reset();
  if (sm_budget != sms) codec_gemms_retune(gemms, sm_budget), sms = sm_budget;
}

void MimiImpl::build_encoder(const GetFn& get, const WeightTable& wt) {
  const int D = spec.dim, ds = spec.downsample_stride;
  DSM_CHECK(ds == 2, DSM_ERR_UNSUPPORTED, "downsample stride %d (only 2: 25 Hz -> 12.5 Hz)", ds);
  d_pcm = arena.alloc<float>((size_t)B * DSM_FRAME_SIZE);
  std::vector<SeanetLayer> L = encoder_layers(spec);
  enc_state0 = arena.alloc<float>((size_t)B * (L[0].k - 1));
  int t_out = DSM_FRAME_SIZE;
  for (auto& y : L) t_out /= y.stride;
  DSM_CHECK(t_out == ds, DSM_ERR_UNSUPPORTED, "SEANet emits %d frames per step, expected %d", t_out, ds);
  // downsample conv input [B, S + T, D] with S = k - stride = ds
  const int Sds = ds, Tds = t_out;
  x_ds = arena.alloc<float>((size_t)B * (Sds + Tds) * D);
  const RowMap x_map{Tds, (long long)(Sds + Tds) * D, (long long)D};
  float* x = x_ds + (long long)Sds * D;
  dbg_seanet = arena.alloc<float>((size_t)B * Tds * D);
  dbg_xfmr = arena.alloc<float>((size_t)B * Tds * D);
  latent = arena.alloc<float>((size_t)B * D);
  resid = arena.alloc<float>((size_t)B * 2 * spec.qdim);
  scores = arena.alloc<float>((size_t)B * spec.bins);

  SeanetBuild sb{this, &arena, B, &get, &wt, &enc_plan, &enc_commits};
  build_seanet(sb, L, DSM_FRAME_SIZE, nullptr, x, x_map, d_pcm, enc_state0, nullptr);

  const int Bc = B;
  const size_t row_bytes = (size_t)D * sizeof(float);
  const size_t pitch = (size_t)(Sds + Tds) * D * sizeof(float);
  float* dbg1 = dbg_seanet;
  float* dbg2 = dbg_xfmr;
  enc_plan.push_back([=](cudaStream_t st) {
    DSM_CUDA(cudaMemcpy2DAsync(dbg1, Tds * row_bytes, x, pitch, Tds * row_bytes, Bc, cudaMemcpyDeviceToDevice, st));
  });
  enc_tr.wsp = wsp;
  enc_tr.init(arena, spec, B, "encoder_transformer", get);
  enc_tr.append_plan(enc_plan, x, x_map, d_mask);
  enc_plan.push_back([=](cudaStream_t st) {
    DSM_CUDA(cudaMemcpy2DAsync(dbg2, Tds * row_bytes, x, pitch, Tds * row_bytes, Bc, cudaMemcpyDeviceToDevice, st));
  });

  // ---- downsample conv (conv.rs:505-534: k = 2*stride, no bias, replicate pad on the first step) ----
  {
    std::vector<float> w = get("downsample.conv.conv.conv.weight", (size_t)D * D * 2 * ds, 1.f / sqrtf((float)D * 2 * ds), 0.f);
    GemmParams g;
    g.A = x_ds, g.a_map = RowMap{1, (long long)(Sds + Tds) * D, (long long)ds * D};
    g.M = Bc, g.N = D, g.K = 2 * ds * D;
    g.out_raw = latent, g.o_map = dense_rows(1, D);
    auto cg = make_codec_gemm(arena, g, conv_to_gemm(w, D, D, 2 * ds), wsp);
    enc_commits.push_back(CommitEntry{x_ds, x_ds + (long long)Tds * D, (long long)(Sds + Tds) * D, (long long)(Sds + Tds) * D, Sds * D, 1});
    float* xds = x_ds;
    int* first = d_first_step;
    enc_plan.push_back([=](cudaStream_t st) {
      launch_replicate_pad(xds, (long long)(Sds + Tds) * D, Sds, D, first, Bc, st);
      run_codec_gemm(*cg, st);
    });
  }
  // ---- split RVQ encode (quantization.rs:366-378) ----
  {
    const int qd = spec.qdim, bins = spec.bins, nq = spec.n_q;
    // both 1x1 input projections in one GEMM: rows [0,qd) = rvq_first, [qd,2qd) = rvq_rest
    std::vector<float> wf = get("quantizer.rvq_first.input_proj.weight", (size_t)qd * D, 1.f / sqrtf((float)D), 0.f);
    std::vector<float> wr = nq > 1 ? get("quantizer.rvq_rest.input_proj.weight", (size_t)qd * D, 1.f / sqrtf((float)D), 0.f)
                                   : std::vector<float>((size_t)qd * D, 0.f);
    wf.insert(wf.end(), wr.begin(), wr.end());
    // One-kernel RVQ (rvq_fused.cu): correct (codes identical) but measured SLOWER than the GEMM + select chain -- codec
    // step 3.36 vs 2.93 ms at 256 streams, equal at 64 (its 128 CTAs pull 32 MB of codebook slices per level through
    // L2) -- so it is opt-in: DSM_RVQ_FUSED_MAX=<largest batch that uses it> (default 0 = never).
    static const int fused_max = [] { const char* e = getenv("DSM_RVQ_FUSED_MAX"); return e ? atoi(e) : 0; }();
    const bool fused = rvq_fused_supported(bins, qd) && Bc <= fused_max;
    GemmParams gp;
    gp.A = latent, gp.a_map = dense_rows(1, D), gp.M = Bc, gp.N = 2 * qd, gp.K = D;
    gp.out_raw = resid, gp.o_map = dense_rows(1, 2 * qd);
    // the residual also as bf16 planes (same [B, 2*qd] row layout): the 32 score GEMMs read it pre-split
    const long long RP = (long long)Bc * 2 * qd;
    __nv_bfloat16* resid_pl = (codec_presplit_enabled() && !fused) ? arena.alloc<__nv_bfloat16>((size_t)3 * RP) : nullptr;
    if (resid_pl) gp.out_planes = resid_pl, gp.p_map = dense_rows(1, 2 * qd), gp.out_plane_stride = RP;
    auto g_proj = make_codec_gemm(arena, gp, wf, wsp);
    if (!g_proj->use_tc) resid_pl = nullptr;     // the CUDA-core GEMM does not write planes
    // codebook embeddings on the host once more for the score GEMMs' weight planes
    std::vector<float> emb_host((size_t)nq * bins * qd);
    DSM_CUDA(cudaMemcpy(emb_host.data(), emb_all, emb_host.size() * sizeof(float), cudaMemcpyDeviceToHost));
    float* embT = nullptr;
    if (fused) {
      // codebooks transposed to [n_q][qdim][bins]: a chunk of dims x the CTA's bins is then a run of rows
      std::vector<float> t(emb_host.size());
      for (int q = 0; q < nq; ++q)
        for (int j = 0; j < bins; ++j)
          for (int d = 0; d < qd; ++d) t[((size_t)q * qd + d) * bins + j] = emb_host[((size_t)q * bins + j) * qd + d];
      embT = arena.upload(t);
    }
    // Grouped encode (rvq_gram.cu): one score GEMM + one argmin-chain kernel per group of DSM_RVQ_GROUP codebooks
    // (default 8; 1 = one GEMM + select per codebook, the reference's association).
    static const int rvq_group = [] {
      const char* e = getenv("DSM_RVQ_GROUP");
      return std::min(std::max(e ? atoi(e) : DSM_RVQ_MAX_GROUP, 1), DSM_RVQ_MAX_GROUP);
    }();
    const bool grouped = !fused && rvq_group > 1 && rvq_chain_supported(bins, rvq_group);
    if (grouped) {
      struct Group { int q0, g, off; std::shared_ptr<CodecGemm> gemm; float* gram; };
      std::vector<Group> groups;
      groups.push_back(Group{0, 1, 0, nullptr, nullptr});                    // rvq_first: its own projected input
      for (int q = 1; q < nq; q += rvq_group) groups.push_back(Group{q, std::min(rvq_group, nq - q), qd, nullptr, nullptr});
      float* gscores = arena.alloc<float>((size_t)Bc * rvq_group * bins);
      for (Group& G : groups) {
        GemmParams s;
        s.A = resid + G.off, s.a_map = RowMap{1, (long long)2 * qd, (long long)2 * qd};
        s.M = Bc, s.N = G.g * bins, s.K = qd;
        s.bias = c2_all + (size_t)G.q0 * bins, s.neg = 1;     // |e|^2/2 - r.e for every codebook of the group
        s.out_raw = gscores, s.o_map = dense_rows(1, G.g * bins);
        if (resid_pl) s.a_planes = resid_pl + G.off, s.a_plane_stride = RP;
        std::vector<float> Eg(emb_host.begin() + (size_t)G.q0 * bins * qd, emb_host.begin() + (size_t)(G.q0 + G.g) * bins * qd);
        G.gemm = make_codec_gemm(arena, s, Eg, wsp);
        if (G.g > 1) {
          const size_t tbl = (size_t)bins * bins;
          G.gram = arena.alloc<float>((size_t)G.g * (G.g - 1) / 2 * tbl, false);
          for (int p = 0; p < G.g; ++p)
            for (int q = p + 1; q < G.g; ++q)
              launch_rvq_gram_table(emb_all + (size_t)(G.q0 + p) * bins * qd, emb_all + (size_t)(G.q0 + q) * bins * qd, bins, qd,
                                    G.gram + (size_t)(p * G.g - p * (p + 1) / 2 + (q - p - 1)) * tbl, stream);
        }
      }
      float* res = resid;
      float* emb = emb_all;
      uint32_t* codes = d_codes;
      enc_plan.push_back([=](cudaStream_t st) {
        run_codec_gemm(*g_proj, st);
        for (const Group& G : groups) {
          run_codec_gemm(*G.gemm, st);
          launch_rvq_chain(gscores, Bc, G.g, bins, G.gram, emb + (size_t)G.q0 * bins * qd, qd, res + G.off, 2 * qd, codes, nq, G.q0,
                           resid_pl ? resid_pl + G.off : nullptr, RP, st);
        }
      });
    }
    std::vector<std::shared_ptr<CodecGemm>> g_score(fused || grouped ? 0 : nq);
    for (int q = 0; q < (fused || grouped ? 0 : nq); ++q) {
      GemmParams s;
      s.A = resid + (q == 0 ? 0 : qd), s.a_map = RowMap{1, (long long)2 * qd, (long long)2 * qd};
      s.M = Bc, s.N = bins, s.K = qd;
      s.bias = c2_all + (size_t)q * bins, s.neg = 1;     // c2 - x.e   (quantization.rs:128-129)
      s.out_raw = scores, s.o_map = dense_rows(1, bins);
      if (resid_pl) s.a_planes = resid_pl + (q == 0 ? 0 : qd), s.a_plane_stride = RP;
      std::vector<float> Eq(emb_host.begin() + (size_t)q * bins * qd, emb_host.begin() + (size_t)(q + 1) * bins * qd);
      g_score[q] = make_codec_gemm(arena, s, Eq, wsp);
    }
    float* res = resid;
    float* sc = scores;
    float* emb = emb_all;
    uint32_t* codes = d_codes;
    const float* c2p = c2_all;
    if (fused) {
      enc_plan.push_back([=](cudaStream_t st) {
        run_codec_gemm(*g_proj, st);
        launch_rvq_fused(embT, c2p, emb, res, Bc, nq, bins, qd, codes, st);
      });
    } else if (!grouped)
    enc_plan.push_back([=](cudaStream_t st) {
      run_codec_gemm(*g_proj, st);
      for (int q = 0; q < nq; ++q) {
        run_codec_gemm(*g_score[q], st);
        launch_rvq_select(sc, Bc, bins, emb + (size_t)q * bins * qd, qd, res + (q == 0 ? 0 : qd), 2 * qd, codes, nq, q, st,
                          resid_pl ? resid_pl + (q == 0 ? 0 : qd) : nullptr, RP);
      }
    });
  }
  // ---- state commit (conv.rs:347-367) ----
  enc_max_count = 0;
  for (auto& e : enc_commits) enc_max_count = std::max(enc_max_count, e.count);
  d_enc_commits = arena.upload(enc_commits);
  {
    CommitEntry* ce = d_enc_commits;
    const int ne = (int)enc_commits.size(), mc = enc_max_count;
    const uint8_t* mk = d_mask;
    int* first = d_first_step;
    enc_plan.push_back([=](cudaStream_t st) { launch_commit_state(ce, ne, mc, mk, first, Bc, st); });
  }
}

void MimiImpl::build_decoder(const GetFn& get, const WeightTable& wt) {
  const int D = spec.dim, ds = spec.downsample_stride, qd = spec.qdim, bins = spec.bins, nq = spec.n_q;
  const int Bc = B;
  d_codes_in = arena.alloc<uint32_t>((size_t)B * nq);
  q_sum = arena.alloc<float>((size_t)B * 2 * qd);
  dec_emb = arena.alloc<float>((size_t)B * D);
  up_state = arena.alloc<float>((size_t)B * D);
  d_pcm_out = arena.alloc<float>((size_t)B * DSM_FRAME_SIZE);
  h_pcm_out = static_cast<float*>(arena.alloc_pinned((size_t)B * DSM_FRAME_SIZE * 4));
  std::vector<SeanetLayer> L = decoder_layers(spec);
  // decoder's first conv input [B, S0 + ds, D]; the decoder transformer stream lives in its last ds rows
  const int S0 = L[0].k - L[0].stride, T0 = ds;
  float* x_d0 = arena.alloc<float>((size_t)B * (S0 + T0) * D);
  float* x = x_d0 + (long long)S0 * D;
  const RowMap x_map{T0, (long long)(S0 + T0) * D, (long long)D};

  // ---- RVQ decode: gather-sum + the two 1x1 output projections ----
  {
    std::vector<float> wf = get("quantizer.rvq_first.output_proj.weight", (size_t)D * qd, 1.f / sqrtf((float)qd), 0.f);
    std::vector<float> wr = nq > 1 ? get("quantizer.rvq_rest.output_proj.weight", (size_t)D * qd, 1.f / sqrtf((float)qd), 0.f)
                                   : std::vector<float>((size_t)D * qd, 0.f);
    std::vector<float> wu = get("upsample.convtr.convtr.convtr.weight", (size_t)D * 2 * ds, 0.7f, 0.f);
    float* dWu = arena.upload(wu);
    uint32_t* cin = d_codes_in;
    float* qs = q_sum;
    float* emb = emb_all;
    float* de = dec_emb;
    float* ups = up_state;
    // dec_emb = Wf . q_first + Wr . q_rest, each projection on its own dense [B, qd] gather buffer
    // (quantization.rs:380-390 decodes the two groups separately and adds)
    GemmParams g1;
    g1.A = q_sum, g1.a_map = dense_rows(1, qd), g1.M = Bc, g1.N = D, g1.K = qd;
    g1.out_raw = dec_emb, g1.o_map = dense_rows(1, D);
    auto cg1 = make_codec_gemm(arena, g1, wf, wsp);
    GemmParams g2;
    g2.A = q_sum + (size_t)Bc * qd, g2.a_map = dense_rows(1, qd), g2.M = Bc, g2.N = D, g2.K = qd;
    g2.res = dec_emb, g2.r_map = dense_rows(1, D), g2.out_raw = dec_emb, g2.o_map = dense_rows(1, D);
    auto cg2 = make_codec_gemm(arena, g2, wr, wsp);
    dec_plan.push_back([=](cudaStream_t st) {
      launch_rvq_gather_sum(cin, nq, 0, 1, emb, bins, qd, Bc, qs, st);
      run_codec_gemm(*cg1, st);
      if (nq > 1) {
        launch_rvq_gather_sum(cin, nq, 1, nq, emb, bins, qd, Bc, qs + (size_t)Bc * qd, st);
        run_codec_gemm(*cg2, st);
      }
      launch_upsample_dw(de, ups, dWu, Bc, D, ds, x, x_map, st);
    });
    dec_commits.push_back(CommitEntry{up_state, dec_emb, (long long)D, (long long)D, D, 0});
  }
  dec_tr.wsp = wsp;
  dec_tr.init(arena, spec, B, "decoder_transformer", get);
  dec_tr.append_plan(dec_plan, x, x_map, d_mask);

  SeanetBuild sb{this, &arena, B, &get, &wt, &dec_plan, &dec_commits};
  build_seanet(sb, L, T0, x_d0, nullptr, RowMap{1, 0, 0}, nullptr, nullptr, d_pcm_out);

  dec_max_count = 0;
  for (auto& e : dec_commits) dec_max_count = std::max(dec_max_count, e.count);
  d_dec_commits = arena.upload(dec_commits);
  {
    CommitEntry* ce = d_dec_commits;
    const int ne = (int)dec_commits.size(), mc = dec_max_count;
    const uint8_t* mk = d_mask;
    dec_plan.push_back([=](cudaStream_t st) { launch_commit_state(ce, ne, mc, mk, nullptr, Bc, st); });
  }
}

void MimiImpl::upload_mask(const uint8_t* host_mask, cudaStream_t st) {
  // a pageable source is staged by the runtime before the call returns, so the caller's buffer (and
  // ours) can be reused immediately even when the step itself stays asynchronous
  if (host_mask) DSM_CUDA(cudaMemcpyAsync(d_mask, host_mask, B, cudaMemcpyHostToDevice, st));
  else DSM_CUDA(cudaMemsetAsync(d_mask, 1, B, st));
}

void MimiImpl::enqueue_encode(const uint8_t* /*mask_dev*/, cudaStream_t st) {
  DSM_CHECK(spec.with_encoder, DSM_ERR_INVALID_ARG, "this Mimi handle was created without an encoder");
  run_plan(enc_plan, st);
}
void MimiImpl::enqueue_decode(const uint8_t* /*mask_dev*/, cudaStream_t st) {
  DSM_CHECK(spec.with_decoder, DSM_ERR_INVALID_ARG, "this Mimi handle was created without a decoder");
  run_plan(dec_plan, st);
}

void MimiImpl::encode_step(const float* pcm, const uint8_t* mask, uint32_t* codes) {
  DSM_CHECK(spec.with_encoder, DSM_ERR_INVALID_ARG, "this Mimi handle was created without an encoder");
  DSM_CUDA(cudaSetDevice(device));
  upload_mask(mask, stream);
  DSM_CUDA(cudaMemcpyAsync(d_pcm, pcm, (size_t)B * DSM_FRAME_SIZE * sizeof(float), cudaMemcpyDefault, stream));
  if (graphs_enabled()) {
    if (!enc_graph.ready()) enc_graph.capture(stream, [&] { run_plan(enc_plan, stream); });
    enc_graph.launch(stream);
  } else {
    run_plan(enc_plan, stream);
  }
  if (codes) {
    const size_t bytes = (size_t)B * spec.n_q * sizeof(uint32_t);
    if (is_device_pointer(codes)) {
      DSM_CUDA(cudaMemcpyAsync(codes, d_codes, bytes, cudaMemcpyDeviceToDevice, stream));
    } else {
      DSM_CUDA(cudaMemcpyAsync(h_codes, d_codes, bytes, cudaMemcpyDeviceToHost, stream));
      DSM_CUDA(cudaStreamSynchronize(stream));
      memcpy(codes, h_codes, bytes);
    }
  }
}

void MimiImpl::decode_step(const uint32_t* codes, const uint8_t* mask, float* pcm) {
  DSM_CHECK(spec.with_decoder, DSM_ERR_INVALID_ARG, "this Mimi handle was created without a decoder");
  DSM_CUDA(cudaSetDevice(device));
  upload_mask(mask, stream);
  DSM_CUDA(cudaMemcpyAsync(d_codes_in, codes, (size_t)B * spec.n_q * sizeof(uint32_t), cudaMemcpyDefault, stream));
  if (graphs_enabled()) {
    if (!dec_graph.ready()) dec_graph.capture(stream, [&] { run_plan(dec_plan, stream); });
    dec_graph.launch(stream);
  } else {
    run_plan(dec_plan, stream);
  }
  if (pcm) {
    const size_t bytes = (size_t)B * DSM_FRAME_SIZE * sizeof(float);
    if (is_device_pointer(pcm)) {
      DSM_CUDA(cudaMemcpyAsync(pcm, d_pcm_out, bytes, cudaMemcpyDeviceToDevice, stream));
    } else {
      DSM_CUDA(cudaMemcpyAsync(h_pcm_out, d_pcm_out, bytes, cudaMemcpyDeviceToHost, stream));
      DSM_CUDA(cudaStreamSynchronize(stream));
      memcpy(pcm, h_pcm_out, bytes);
    }
  }
}

void MimiImpl::reset_state() {
  // mimi.rs:227-234
  DSM_CUDA(cudaSetDevice(device));
  const int one = 1;
  DSM_CUDA(cudaMemcpyAsync(d_first_step, &one, sizeof(int), cudaMemcpyHostToDevice, stream));
  DSM_CUDA(cudaMemcpyAsync(d_first_step_dec, &one, sizeof(int), cudaMemcpyHostToDevice, stream));
  DSM_CUDA(cudaStreamSynchronize(stream));
  if (d_enc_commits) launch_zero_rows(d_enc_commits, (int)enc_commits.size(), 0, B, stream);
  if (d_dec_commits) launch_zero_rows(d_dec_commits, (int)dec_commits.size(), 0, B, stream);
  if (spec.with_encoder) enc_tr.reset_all(stream);
  if (spec.with_decoder) dec_tr.reset_all(stream);
  DSM_CUDA(cudaStreamSynchronize(stream));
}

void MimiImpl::reset_batch_idx(int b) {
  // mimi.rs:236-244 (and the decoder transformer, which the reference forgets -- SURVEY.md 3.5)
  DSM_CHECK(b >= 0 && b < B, DSM_ERR_INVALID_ARG, "batch_idx %d out of range [0, %d)", b, B);
  DSM_CUDA(cudaSetDevice(device));
  if (d_enc_commits) launch_zero_rows(d_enc_commits, (int)enc_commits.size(), b, 1, stream);
  if (d_dec_commits) launch_zero_rows(d_dec_commits, (int)dec_commits.size(), b, 1, stream);
  if (spec.with_encoder) enc_tr.reset_batch_idx(b, stream);
  if (spec.with_decoder) dec_tr.reset_batch_idx(b, stream);
}

}  // namespace dsm
