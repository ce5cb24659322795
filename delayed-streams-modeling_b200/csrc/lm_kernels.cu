// Hand-written kernels of the delayed-streams LM step that are not GEMMs: embedding sum, RMSNorm,
// RoPE + ring scatter, single-query ring attention (bf16 KV, f32 math), SiLU gate, greedy argmax,
// VAD extra heads, and the ASR token feed / commit.
#include <algorithm>

#include "kernels.cuh"

namespace dsm {

// ================================================================================================
// emb = E_text[t] + sum_i E_audio_i[c_i] (+ cond)   (lm.rs:983-1000).  f32 accumulation, in the
// reference's order.  One CTA per stream, 8 contiguous bf16 per thread-load.
// ================================================================================================
__global__ void __launch_bounds__(256) embed_sum_kernel(const uint32_t* __restrict__ text_ids,
                                                         const uint32_t* __restrict__ audio_ids, int n_cb,
                                                         const __nv_bfloat16* __restrict__ text_emb, int text_vocab,
                                                         const __nv_bfloat16* __restrict__ audio_emb, int audio_vocab,
                                                         const float* __restrict__ cond, int D, float* __restrict__ x) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  uint32_t tid_text = text_ids[b];
  if (tid_text >= (uint32_t)text_vocab) tid_text = text_vocab - 1;
  for (int c = threadIdx.x * 8; c < D; c += blockDim.x * 8) {
    float acc[8];
    {
      uint4 u = __ldg(reinterpret_cast<const uint4*>(text_emb + (long long)tid_text * D + c));
      acc[0] = bf16_lo(u.x), acc[1] = bf16_hi(u.x), acc[2] = bf16_lo(u.y), acc[3] = bf16_hi(u.y);
      acc[4] = bf16_lo(u.z), acc[5] = bf16_hi(u.z), acc[6] = bf16_lo(u.w), acc[7] = bf16_hi(u.w);
    }
    for (int i = 0; i < n_cb; ++i) {
      uint32_t id = audio_ids[(long long)b * n_cb + i];
      if (id >= (uint32_t)audio_vocab) id = audio_vocab - 1;
      uint4 u = __ldg(reinterpret_cast<const uint4*>(audio_emb + ((long long)i * audio_vocab + id) * D + c));
      acc[0] += bf16_lo(u.x), acc[1] += bf16_hi(u.x), acc[2] += bf16_lo(u.y), acc[3] += bf16_hi(u.y);
      acc[4] += bf16_lo(u.z), acc[5] += bf16_hi(u.z), acc[6] += bf16_lo(u.w), acc[7] += bf16_hi(u.w);
    }
    if (cond) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += cond[c + j];
    }
    float* o = x + (long long)b * D + c;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

void launch_embed_sum(const uint32_t* text_ids, const uint32_t* audio_ids, int n_cb, const __nv_bfloat16* text_emb,
                      int text_vocab, const __nv_bfloat16* audio_emb, int audio_vocab, const float* cond, int B, int D,
                      float* x, cudaStream_t st) {
  DSM_CHECK(D % 8 == 0, DSM_ERR_UNSUPPORTED, "embed_sum: D=%d", D);
  launch_k(embed_sum_kernel, B, 256, 0, st, text_ids, audio_ids, n_cb, text_emb, text_vocab, audio_emb, audio_vocab, cond, D, x);
  check_launch("embed_sum");
}

// ================================================================================================
// RMSNorm (batched_transformer.rs:181-198, eps 1e-8): x * rsqrt(mean(x^2) + eps) * alpha, f32 math,
// bf16 (GEMM operand) and optional f32 outputs.  One CTA per stream.
// ================================================================================================
constexpr int kMaxPartials = 16;   // == TC_MAX_SPLIT (gemm_tc.cu)

// Split-K partial sums of the projection that precedes the norm ([S][B][D] planes, see gemm_tc.cuh) are folded into
// the residual stream here, in split order: x <- x + (P_0 + P_1 + ... + P_{S-1}).
__global__ void __launch_bounds__(256) rmsnorm_kernel(float* __restrict__ x, const float* __restrict__ partials, int S,
                                                       long long part_stride, const float* __restrict__ alpha, float eps,
                                                       int D, __nv_bfloat16* __restrict__ out_bf16,
                                                       float* __restrict__ out_f32) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_part[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  float* row = x + (long long)b * D;
  float ss = 0.f;
  for (int c = tid * 4; c < D; c += blockDim.x * 4) {
    float4 v = *reinterpret_cast<const float4*>(row + c);
    if (S > 0) {
      // all planes requested before the first add (the loop bound is a run-time value: unrolled by hand to 16)
      float4 q[kMaxPartials];
#pragma unroll
      for (int sp = 0; sp < kMaxPartials; ++sp)
        if (sp < S) q[sp] = __ldcg(reinterpret_cast<const float4*>(partials + sp * part_stride + (long long)b * D + c));
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int sp = 0; sp < kMaxPartials; ++sp)
        if (sp < S) acc.x += q[sp].x, acc.y += q[sp].y, acc.z += q[sp].z, acc.w += q[sp].w;
      v.x += acc.x, v.y += acc.y, v.z += acc.z, v.w += acc.w;
      *reinterpret_cast<float4*>(row + c) = v;
    }
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) s_part[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += s_part[w];
  const float r = rsqrtf(tot / (float)D + eps);
  for (int c = tid * 4; c < D; c += blockDim.x * 4) {
    float4 v = *reinterpret_cast<const float4*>(row + c);      // this thread's own stores above
    float4 a = *reinterpret_cast<const float4*>(alpha + c);
    float4 o = make_float4(v.x * r * a.x, v.y * r * a.y, v.z * r * a.z, v.w * r * a.w);
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + (long long)b * D + c) = o;
    if (out_bf16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo);
      u.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(out_bf16 + (long long)b * D + c) = u;
    }
  }
}

void launch_rmsnorm(float* x, const PartialSums& ps, const float* alpha, float eps, int B, int D, __nv_bfloat16* out_bf16,
                    float* out_f32, cudaStream_t st) {
  DSM_CHECK(D % 4 == 0 && ps.count <= kMaxPartials, DSM_ERR_UNSUPPORTED, "rmsnorm: D=%d, %d partial planes", D, ps.count);
  launch_k(rmsnorm_kernel, B, 256, 0, st, x, ps.planes, ps.count, ps.stride, alpha, eps, D, out_bf16, out_f32);
  check_launch("rmsnorm");
}

// ================================================================================================
// RoPE in f32 (transformer.rs:372-378) on q, k; q -> bf16 [B, H*D]; k, v -> bf16 ring slot cur_index[b].
// ================================================================================================
// The QKV projection may arrive as split-K partial planes [S][B][3*H*D] (S > 0): they are added here, in split order.
__device__ __forceinline__ float2 load_qkv2(const float* __restrict__ qkv, int S, long long stride, long long off) {
  if (S == 0) return *reinterpret_cast<const float2*>(qkv + off);
  float2 v[kMaxPartials];
#pragma unroll
  for (int sp = 0; sp < kMaxPartials; ++sp)
    if (sp < S) v[sp] = __ldcg(reinterpret_cast<const float2*>(qkv + sp * stride + off));
  float2 a = make_float2(0.f, 0.f);
#pragma unroll
  for (int sp = 0; sp < kMaxPartials; ++sp)
    if (sp < S) a.x += v[sp].x, a.y += v[sp].y;
  return a;
}

__global__ void __launch_bounds__(256) rope_scatter_bf16_kernel(const float* __restrict__ qkv, int S, long long part_stride,
                                                                 RingState rs, int H,
                                                                 int D, int context, const float* __restrict__ inv_freq,
                                                                 __nv_bfloat16* __restrict__ q_out,
                                                                 __nv_bfloat16* __restrict__ k_ring,
                                                                 __nv_bfloat16* __restrict__ v_ring) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int slot = (int)rs.cur_index[b];
  const float pos = (float)rs.rope_pos[b];
  const int HD = H * D, half = D / 2;
  const long long row = (long long)b * 3 * HD;
  for (int i = threadIdx.x; i < H * half; i += blockDim.x) {
    const int h = i / half, p = i - h * half;
    float sn, cs;
    sincosf(pos * inv_freq[p], &sn, &cs);
    const float2 q = load_qkv2(qkv, S, part_stride, row + h * D + 2 * p);
    const float2 k = load_qkv2(qkv, S, part_stride, row + HD + h * D + 2 * p);
    const float2 v = load_qkv2(qkv, S, part_stride, row + 2 * HD + h * D + 2 * p);
    *reinterpret_cast<__nv_bfloat162*>(q_out + (long long)b * HD + h * D + 2 * p) =
        __floats2bfloat162_rn(q.x * cs - q.y * sn, q.x * sn + q.y * cs);
    const long long o = (((long long)b * H + h) * context + slot) * D + 2 * p;
    *reinterpret_cast<__nv_bfloat162*>(k_ring + o) = __floats2bfloat162_rn(k.x * cs - k.y * sn, k.x * sn + k.y * cs);
    *reinterpret_cast<__nv_bfloat162*>(v_ring + o) = __floats2bfloat162_rn(v.x, v.y);
  }
}

void launch_rope_scatter_bf16(const float* qkv, const PartialSums& ps, const RingState& rs, int B, int H, int D, int context,
                              const float* inv_freq, __nv_bfloat16* q_out, __nv_bfloat16* k_ring,
                              __nv_bfloat16* v_ring, cudaStream_t st) {
  launch_k(rope_scatter_bf16_kernel, B, 256, 0, st, ps.count ? ps.planes : qkv, ps.count, ps.stride, rs, H, D, context, inv_freq,
           q_out, k_ring, v_ring);
  check_launch("rope_scatter_bf16");
}

// ================================================================================================
// Single-query attention over the per-stream ring (batched_transformer.rs:107-113, mask of
// kv_cache.rs:130-170: slot j visible iff j <= position before the advance).  THE dominant kernel of
// the step: per (stream, head) it streams ctx*D*2 bf16 of K and of V from HBM exactly once.
// One CTA (4 warps) per (stream, head).  A key row of D bf16 is covered by D/8 lanes x 16 B, so one
// warp-wide load instruction fetches 32*16 B = 512 contiguous bytes (256/D keys... 2 keys at D=128,
// 4 at D=64); loads are unrolled UNROLL deep to keep >= 4 KB per warp in flight.
// ================================================================================================
// 16-byte streaming load that does not allocate in L1 (the rings are read exactly once per step), so the
// kernel's bandwidth does not depend on how much of the SM's L1/shared array is carved out as L1.
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
template <int D, int UNROLL>
__global__ void __launch_bounds__(128, 8) lm_attention_kernel(const __nv_bfloat16* __restrict__ q, RingState rs,
                                                               const uint8_t* __restrict__ mask, int H, int context,
                                                               const __nv_bfloat16* __restrict__ k_ring,
                                                               const __nv_bfloat16* __restrict__ v_ring,
                                                               __nv_bfloat16* __restrict__ out) {
  constexpr int LPK = D / 8;          // lanes per key
  constexpr int KPW = 32 / LPK;       // keys per warp-load
  constexpr int NW = 4;
  extern __shared__ float sm[];
  float* sc = sm;                     // [context] scores, then reused as [NW][D] partial outputs
  float* red = sc;
  __shared__ float s_red[NW];
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  auto group_sync = [&]() { __syncthreads(); };
  const int HD = H * D;
  pdl_launch_dependents();
  pdl_wait();
  {
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  __nv_bfloat16* orow = out + (long long)b * HD + h * D;
  const bool active = mask ? mask[b] != 0 : true;
  if (!active) {
    // an inactive stream's output is never consumed (asr.rs:221-223); skip the ring read
    for (int d = tid; d < D; d += 128) orow[d] = __float2bfloat16_rn(0.f);
    return;
  }
  const int n = ring_visible_count(rs.cur_pos[b], context);             // visible slots are [0, n)
  const int c8 = lane % LPK, sub = lane / LPK;
  const float scale = rsqrtf((float)D);

  float qf[8];
  {
    uint4 u = *reinterpret_cast<const uint4*>(q + (long long)b * HD + h * D + c8 * 8);
    qf[0] = bf16_lo(u.x) * scale, qf[1] = bf16_hi(u.x) * scale, qf[2] = bf16_lo(u.y) * scale, qf[3] = bf16_hi(u.y) * scale;
    qf[4] = bf16_lo(u.z) * scale, qf[5] = bf16_hi(u.z) * scale, qf[6] = bf16_lo(u.w) * scale, qf[7] = bf16_hi(u.w) * scale;
  }
  const __nv_bfloat16* kb = k_ring + ((long long)b * H + h) * context * D;
  const __nv_bfloat16* vb = v_ring + ((long long)b * H + h) * context * D;

  // ---- scores ----
  for (int j0 = warp * KPW * UNROLL; j0 < n; j0 += NW * KPW * UNROLL) {
    uint4 kv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int j = j0 + u * KPW + sub;
      kv[u] = j < n ? ld_stream16(kb + (long long)j * D + c8 * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      float s = qf[0] * bf16_lo(kv[u].x);
      s = fmaf(qf[1], bf16_hi(kv[u].x), s);
      s = fmaf(qf[2], bf16_lo(kv[u].y), s);
      s = fmaf(qf[3], bf16_hi(kv[u].y), s);
      s = fmaf(qf[4], bf16_lo(kv[u].z), s);
      s = fmaf(qf[5], bf16_hi(kv[u].z), s);
      s = fmaf(qf[6], bf16_lo(kv[u].w), s);
      s = fmaf(qf[7], bf16_hi(kv[u].w), s);
#pragma unroll
      for (int o = LPK / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const int j = j0 + u * KPW + sub;
      if (c8 == 0 && j < n) sc[j] = s;
    }
  }
  group_sync();

  // ---- softmax (f32) ----
  float mx = -INFINITY;
  for (int j = tid; j < n; j += 128) mx = fmaxf(mx, sc[j]);
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  group_sync();
  mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  group_sync();
  float sum = 0.f;
  for (int j = tid; j < n; j += 128) {
    const float e = __expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  group_sync();
  const float inv = 1.f / (s_red[0] + s_red[1] + s_red[2] + s_red[3]);

  // ---- P.V ----
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int j0 = warp * KPW * UNROLL; j0 < n; j0 += NW * KPW * UNROLL) {
    uint4 vv[UNROLL];
    float pj[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int j = j0 + u * KPW + sub;
      const bool ok = j < n;
      vv[u] = ok ? ld_stream16(vb + (long long)j * D + c8 * 8) : make_uint4(0, 0, 0, 0);
      pj[u] = ok ? sc[j] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      acc[0] = fmaf(pj[u], bf16_lo(vv[u].x), acc[0]);
      acc[1] = fmaf(pj[u], bf16_hi(vv[u].x), acc[1]);
      acc[2] = fmaf(pj[u], bf16_lo(vv[u].y), acc[2]);
      acc[3] = fmaf(pj[u], bf16_hi(vv[u].y), acc[3]);
      acc[4] = fmaf(pj[u], bf16_lo(vv[u].z), acc[4]);
      acc[5] = fmaf(pj[u], bf16_hi(vv[u].z), acc[5]);
      acc[6] = fmaf(pj[u], bf16_lo(vv[u].w), acc[6]);
      acc[7] = fmaf(pj[u], bf16_hi(vv[u].w), acc[7]);
    }
  }
  // fold the KPW key sub-groups of a warp, then the warps
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = LPK; o < 32; o <<= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  }
  group_sync();                    // every warp is done reading the probabilities in sc[]
  if (sub == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp * D + c8 * 8 + i] = acc[i];
  }
  group_sync();
  for (int d = tid; d < D; d += 128) {
    const float v = (red[d] + red[D + d] + red[2 * D + d] + red[3 * D + d]) * inv;
    orow[d] = __float2bfloat16_rn(v);
  }
  }
}

void launch_lm_attention(const __nv_bfloat16* q, const RingState& rs, const uint8_t* mask, int B, int H, int D,
                         int context, const __nv_bfloat16* k_ring, const __nv_bfloat16* v_ring, __nv_bfloat16* out,
                         cudaStream_t st) {
  const size_t smem = (size_t)std::max(context, 4 * D) * sizeof(float);
  if (D == 128) launch_k(lm_attention_kernel<128, 8>, B * H, 128, smem, st, q, rs, mask, H, context, k_ring, v_ring, out);
  else if (D == 64) launch_k(lm_attention_kernel<64, 8>, B * H, 128, smem, st, q, rs, mask, H, context, k_ring, v_ring, out);
  else fail(DSM_ERR_UNSUPPORTED, "lm_attention: head_dim %d (64 or 128)", D);
  check_launch("lm_attention");
}

// ================================================================================================
// 8-bit KV ring (opt-in, dsm_lm_options.kv_dtype = DSM_KV_INT8).  The attention above already runs at the
// HBM copy rate, so only fewer bytes make it faster -- and a bf16 ring (98.3 MB per stt-1b stream) caps a 180 GB
// GPU at ~1,700 streams.  Storage per (stream, head):
//   K  int8 [ctx][D]            one key row = D contiguous bytes, scale k_scale[slot] = absmax / 127
//   V  int8 [ctxp/4][D][4]      key-quad interleaved: the 4 bytes of a 32-bit word are the SAME channel d of 4
//                               consecutive slots, scale v_scale[slot]                  (ctxp = ctx rounded up to 4)
// so that both contractions are integer dot products (IDP4A) with the per-slot scales applied to scalars:
//   score_j = kscale_j * qscale * sum_d q8[d] * k8[j][d]          (the reduction over d runs inside a K word)
//   out[d]  = pscale   * sum_j p8[j] * v8[j][d],  p8 = quantised (softmax_j * vscale_j)   (over j inside a V word)
// The query and the probabilities are split into two int8 terms (x ~ s * (x1 + x2 / 254), 15 bits), so the only
// rounding that matters is the ring's own 8 bits.
// ================================================================================================
// One warp per head, a lane owns 4 consecutive channels (two RoPE pairs): q / k / v arrive as one float4 per lane, K
// leaves as one 32-bit store, and V -- whose bytes for this slot sit at stride 4 in the key-quad layout -- as a
// read-modify-write of the lane's four words (one coalesced 512-byte row per head each way; nobody else writes the
// ring, and a (stream, head) belongs to one warp).
__device__ __forceinline__ float4 load_qkv4(const float* __restrict__ qkv, int S, long long stride, long long off) {
  if (S == 0) return *reinterpret_cast<const float4*>(qkv + off);
  float4 v[kMaxPartials];
#pragma unroll
  for (int sp = 0; sp < kMaxPartials; ++sp)
    if (sp < S) v[sp] = __ldcg(reinterpret_cast<const float4*>(qkv + sp * stride + off));
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int sp = 0; sp < kMaxPartials; ++sp)
    if (sp < S) a.x += v[sp].x, a.y += v[sp].y, a.z += v[sp].z, a.w += v[sp].w;
  return a;
}

__global__ void __launch_bounds__(256) rope_scatter_q8_kernel(const float* __restrict__ qkv, int S, long long part_stride,
                                                               RingState rs, int H, int D,
                                                               int context, int ctxp, const float* __restrict__ inv_freq,
                                                               __nv_bfloat16* __restrict__ q_out, int8_t* __restrict__ k_ring,
                                                               int8_t* __restrict__ v_ring, float* __restrict__ k_scale,
                                                               float* __restrict__ v_scale,
                                                               __nv_bfloat16* __restrict__ v_ring_bf16) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int slot = (int)rs.cur_index[b];
  const float pos = (float)rs.rope_pos[b];
  const int HD = H * D;
  const long long row = (long long)b * 3 * HD;
  const int d0 = 4 * lane;                       // this lane's channels [d0, d0 + 4) = pairs 2*lane, 2*lane + 1
  const bool on = d0 < D;
  float sn0 = 0.f, cs0 = 1.f, sn1 = 0.f, cs1 = 1.f;
  if (on) {
    sincosf(pos * inv_freq[2 * lane], &sn0, &cs0);
    sincosf(pos * inv_freq[2 * lane + 1], &sn1, &cs1);
  }
  for (int h = warp; h < H; h += blockDim.x / 32) {
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f), k = q, v = q;
    if (on) {
      q = load_qkv4(qkv, S, part_stride, row + h * D + d0);
      k = load_qkv4(qkv, S, part_stride, row + HD + h * D + d0);
      v = load_qkv4(qkv, S, part_stride, row + 2 * HD + h * D + d0);
    }
    const float kr[4] = {k.x * cs0 - k.y * sn0, k.x * sn0 + k.y * cs0, k.z * cs1 - k.w * sn1, k.z * sn1 + k.w * cs1};
    const float vr[4] = {v.x, v.y, v.z, v.w};
    float kmax = fmaxf(fmaxf(fabsf(kr[0]), fabsf(kr[1])), fmaxf(fabsf(kr[2]), fabsf(kr[3])));
    float vmax = fmaxf(fmaxf(fabsf(vr[0]), fabsf(vr[1])), fmaxf(fabsf(vr[2]), fabsf(vr[3])));
    kmax = warp_max(kmax), vmax = warp_max(vmax);
    const float kinv = kmax > 0.f ? 127.f / kmax : 0.f, vinv = vmax > 0.f ? 127.f / vmax : 0.f;
    const long long bh = (long long)b * H + h;
    if (lane == 0) {
      k_scale[bh * ctxp + slot] = kmax / 127.f;
      if (!v_ring_bf16) v_scale[bh * ctxp + slot] = vmax / 127.f;
    }
    if (!on) continue;
    {
      __nv_bfloat162 lo = __floats2bfloat162_rn(q.x * cs0 - q.y * sn0, q.x * sn0 + q.y * cs0);
      __nv_bfloat162 hi = __floats2bfloat162_rn(q.z * cs1 - q.w * sn1, q.z * sn1 + q.w * cs1);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo), u.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(q_out + (long long)b * HD + h * D + d0) = u;
    }
    int kq[4], vq[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) kq[i] = __float2int_rn(kr[i] * kinv), vq[i] = __float2int_rn(vr[i] * vinv);
    *reinterpret_cast<uint32_t*>(k_ring + (bh * context + slot) * D + d0) =
        (uint32_t)(kq[0] & 0xff) | ((uint32_t)(kq[1] & 0xff) << 8) | ((uint32_t)(kq[2] & 0xff) << 16) | ((uint32_t)(kq[3] & 0xff) << 24);
    if (v_ring_bf16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(vr[0], vr[1]), hi = __floats2bfloat162_rn(vr[2], vr[3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo), u.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(v_ring_bf16 + (bh * context + slot) * D + d0) = u;
    } else {
      // word d of key quad slot/4 holds channel d of slots 4*(slot/4) .. +3: replace byte slot%4 of four words
      uint4* vw = reinterpret_cast<uint4*>(v_ring + bh * ctxp * D + (long long)(slot >> 2) * D * 4 + d0 * 4);
      uint4 w = *vw;
      const int sh = 8 * (slot & 3);
      const uint32_t keep = ~(0xffu << sh);
      w.x = (w.x & keep) | ((uint32_t)(vq[0] & 0xff) << sh);
      w.y = (w.y & keep) | ((uint32_t)(vq[1] & 0xff) << sh);
      w.z = (w.z & keep) | ((uint32_t)(vq[2] & 0xff) << sh);
      w.w = (w.w & keep) | ((uint32_t)(vq[3] & 0xff) << sh);
      *vw = w;
    }
  }
}

void launch_rope_scatter_q8(const float* qkv, const PartialSums& ps, const RingState& rs, int B, int H, int D, int context,
                            const float* inv_freq, __nv_bfloat16* q_out, const KvRingQ8& r, cudaStream_t st) {
  DSM_CHECK(D == 64 || D == 128, DSM_ERR_UNSUPPORTED, "rope_scatter_q8: head_dim %d (64 or 128)", D);
  launch_k(rope_scatter_q8_kernel, B, 256, 0, st, ps.count ? ps.planes : qkv, ps.count, ps.stride, rs, H, D, context, r.ctxp, inv_freq, q_out, r.k, r.v, r.k_scale, r.v_scale,
                                            r.v_bf16);
  check_launch("rope_scatter_q8");
}

// x -> two int8 terms: x * inv ~ x1 + x2 / 254, |x1| <= 127, |x2| <= 127
__device__ __forceinline__ void split2_i8(float t, int& x1, int& x2) {
  const float r = rintf(t);
  x1 = (int)r;
  x2 = __float2int_rn((t - r) * 254.f);
}
__device__ __forceinline__ uint32_t pack4_i8(int a, int b, int c, int d) {
  return (uint32_t)(a & 0xff) | ((uint32_t)(b & 0xff) << 8) | ((uint32_t)(c & 0xff) << 16) | ((uint32_t)(d & 0xff) << 24);
}

// One CTA (4 warps) per (stream, head), as the bf16 kernel.  K pass: a key row is D/16 lanes x 16 B, 32 / (D/16)
// keys per warp-load.  V pass: a key-quad block is D*4 bytes = D/4 lanes x 16 B; a lane owns 4 channels.
// V8 = false: the V ring stays bf16 [ctx][D] (DSM_KV_INT8K) and P.V is the bf16 kernel's f32 pass.
template <int D, int UNROLL, bool V8>
__global__ void __launch_bounds__(128, 8) lm_attention_q8_kernel(const __nv_bfloat16* __restrict__ q, RingState rs,
                                                                  const uint8_t* __restrict__ mask, int H, int context,
                                                                  int ctxp, const int8_t* __restrict__ k_ring,
                                                                  const int8_t* __restrict__ v_ring,
                                                                  const float* __restrict__ k_scale,
                                                                  const float* __restrict__ v_scale,
                                                                  const __nv_bfloat16* __restrict__ v_ring_bf16,
                                                                  __nv_bfloat16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int LPK = D / 16;         // lanes per key row
  constexpr int KPW = 32 / LPK;       // keys per warp-load
  constexpr int LPB = D / 4;          // lanes per key-quad block
  constexpr int BPW = 32 / LPB;       // blocks per warp-load
  constexpr int NW = 4;
  extern __shared__ float sm[];
  float* sc = sm;                                            // [ctxp] scores, then p * vscale
  uint2* pq = reinterpret_cast<uint2*>(sm + ctxp);           // [ctxp/4] {4 x p1, 4 x p2} per key quad
  int* red = reinterpret_cast<int*>(sm + ctxp + ctxp / 2);   // [NW][D][2] partial integer outputs
  __shared__ float s_red[2][NW];
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int HD = H * D;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  __nv_bfloat16* orow = out + (long long)b * HD + h * D;
  if (mask && !mask[b]) {
    for (int d = tid; d < D; d += 128) orow[d] = __float2bfloat16_rn(0.f);
    return;
  }
  const int n = ring_visible_count(rs.cur_pos[b], context);             // visible slots are [0, n)
  const int n4 = (n + 3) & ~3;
  const long long bh = (long long)b * H + h;
  const float* ks = k_scale + bh * ctxp;
  const float* vs = v_scale + bh * ctxp;

  // ---- query: bf16 -> two int8 terms with one scale per head ----
  const int c16 = lane % LPK, sub = lane / LPK;
  uint32_t q1[4], q2[4];
  float qs;
  {
    const uint4* qp = reinterpret_cast<const uint4*>(q + (long long)b * HD + h * D + c16 * 16);
    const uint4 u0 = qp[0], u1 = qp[1];
    const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
    float qf[16];
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      qf[2 * i] = bf16_lo(w[i]), qf[2 * i + 1] = bf16_hi(w[i]);
      amax = fmaxf(amax, fmaxf(fabsf(qf[2 * i]), fabsf(qf[2 * i + 1])));
    }
#pragma unroll
    for (int o = LPK / 2; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const float inv = amax > 0.f ? 127.f / amax : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int a1[4], a2[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split2_i8(qf[4 * i + e] * inv, a1[e], a2[e]);
      q1[i] = pack4_i8(a1[0], a1[1], a1[2], a1[3]);
      q2[i] = pack4_i8(a2[0], a2[1], a2[2], a2[3]);
    }
    qs = amax * (1.f / (127.f * 254.f)) * rsqrtf((float)D);
  }
  const int8_t* kb = k_ring + bh * context * D;
  const int8_t* vb = v_ring + bh * ctxp * D;

  // ---- scores (integer dot products; the key's scale is applied in the softmax pass) ----
  for (int j0 = warp * KPW * UNROLL; j0 < n; j0 += NW * KPW * UNROLL) {
    uint4 kv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int j = j0 + u * KPW + sub;
      kv[u] = j < n ? ld_stream16(kb + (long long)j * D + c16 * 16) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      int hi = __dp4a((int)kv[u].x, (int)q1[0], 0), lo = __dp4a((int)kv[u].x, (int)q2[0], 0);
      hi = __dp4a((int)kv[u].y, (int)q1[1], hi), lo = __dp4a((int)kv[u].y, (int)q2[1], lo);
      hi = __dp4a((int)kv[u].z, (int)q1[2], hi), lo = __dp4a((int)kv[u].z, (int)q2[2], lo);
      hi = __dp4a((int)kv[u].w, (int)q1[3], hi), lo = __dp4a((int)kv[u].w, (int)q2[3], lo);
      int t = hi * 254 + lo;      // |t| <= 16 * 127 * 127 * 255 < 2^26
#pragma unroll
      for (int o = LPK / 2; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      const int j = j0 + u * KPW + sub;
      if (c16 == 0 && j < n) sc[j] = (float)t;
    }
  }
  __syncthreads();

  // ---- softmax (f32); p_j * vscale_j quantised to two int8 terms against its maximum ----
  float mx = -INFINITY;
  for (int j = tid; j < n; j += 128) {
    const float s = sc[j] * (ks[j] * qs);
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if (lane == 0) s_red[0][warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(s_red[0][0], s_red[0][1]), fmaxf(s_red[0][2], s_red[0][3]));
  __syncthreads();
  float sum = 0.f, pm = 0.f;
  for (int j = tid; j < n; j += 128) {
    const float e = __expf(sc[j] - mx);
    const float pv = V8 ? e * vs[j] : e;
    sc[j] = pv;
    sum += e;
    pm = fmaxf(pm, pv);
  }
  sum = warp_sum(sum), pm = warp_max(pm);
  if (lane == 0) s_red[0][warp] = sum, s_red[1][warp] = pm;
  __syncthreads();
  sum = s_red[0][0] + s_red[0][1] + s_red[0][2] + s_red[0][3];
  if (!V8) {
    // ---- P.V with a bf16 V ring: the bf16 kernel's pass (D/8 lanes x 16 B per key row) ----
    constexpr int LPK2 = D / 8, KPW2 = 32 / LPK2;
    const int c8 = lane % LPK2, sub2 = lane / LPK2;
    const __nv_bfloat16* vbf = v_ring_bf16 + bh * context * D;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int j0 = warp * KPW2 * UNROLL; j0 < n; j0 += NW * KPW2 * UNROLL) {
      uint4 vv[UNROLL];
      float pj[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int j = j0 + u * KPW2 + sub2;
        const bool ok = j < n;
        vv[u] = ok ? ld_stream16(vbf + (long long)j * D + c8 * 8) : make_uint4(0, 0, 0, 0);
        pj[u] = ok ? sc[j] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        acc[0] = fmaf(pj[u], bf16_lo(vv[u].x), acc[0]);
        acc[1] = fmaf(pj[u], bf16_hi(vv[u].x), acc[1]);
        acc[2] = fmaf(pj[u], bf16_lo(vv[u].y), acc[2]);
        acc[3] = fmaf(pj[u], bf16_hi(vv[u].y), acc[3]);
        acc[4] = fmaf(pj[u], bf16_lo(vv[u].z), acc[4]);
        acc[5] = fmaf(pj[u], bf16_hi(vv[u].z), acc[5]);
        acc[6] = fmaf(pj[u], bf16_lo(vv[u].w), acc[6]);
        acc[7] = fmaf(pj[u], bf16_hi(vv[u].w), acc[7]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int o = LPK2; o < 32; o <<= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    __syncthreads();                    // every warp is done reading the probabilities in sc[]
    float* redf = sc;                   // [NW][D], ctxp >= 4 * D is not guaranteed: use the integer scratch when short
    if (ctxp < NW * D) redf = reinterpret_cast<float*>(red);
    if (sub2 == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) redf[warp * D + c8 * 8 + i] = acc[i];
    }
    __syncthreads();
    const float inv = 1.f / sum;
    for (int d = tid; d < D; d += 128) orow[d] = __float2bfloat16_rn((redf[d] + redf[D + d] + redf[2 * D + d] + redf[3 * D + d]) * inv);
    return;
  }
  pm = fmaxf(fmaxf(s_red[1][0], s_red[1][1]), fmaxf(s_red[1][2], s_red[1][3]));
  const float pinv = pm > 0.f ? 127.f / pm : 0.f;
  int8_t* pq8 = reinterpret_cast<int8_t*>(pq);
  for (int j = tid; j < n4; j += 128) {
    int p1 = 0, p2 = 0;
    if (j < n) split2_i8(sc[j] * pinv, p1, p2);
    pq8[(j >> 2) * 8 + (j & 3)] = (int8_t)p1;
    pq8[(j >> 2) * 8 + 4 + (j & 3)] = (int8_t)p2;
  }
  __syncthreads();

  // ---- P.V: per lane 4 channels x {hi, lo} integer accumulators ----
  const int nb = n4 >> 2;
  const int lb = lane % LPB, bsub = lane / LPB;
  int ah[4] = {0, 0, 0, 0}, al[4] = {0, 0, 0, 0};
  for (int kb0 = warp * BPW * UNROLL; kb0 < nb; kb0 += NW * BPW * UNROLL) {
    uint4 vv[UNROLL];
    uint2 pp[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int kq = kb0 + u * BPW + bsub;
      const bool ok = kq < nb;
      vv[u] = ok ? ld_stream16(vb + (long long)kq * D * 4 + lb * 16) : make_uint4(0, 0, 0, 0);
      pp[u] = ok ? pq[kq] : make_uint2(0, 0);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      ah[0] = __dp4a((int)vv[u].x, (int)pp[u].x, ah[0]), al[0] = __dp4a((int)vv[u].x, (int)pp[u].y, al[0]);
      ah[1] = __dp4a((int)vv[u].y, (int)pp[u].x, ah[1]), al[1] = __dp4a((int)vv[u].y, (int)pp[u].y, al[1]);
      ah[2] = __dp4a((int)vv[u].z, (int)pp[u].x, ah[2]), al[2] = __dp4a((int)vv[u].z, (int)pp[u].y, al[2]);
      ah[3] = __dp4a((int)vv[u].w, (int)pp[u].x, ah[3]), al[3] = __dp4a((int)vv[u].w, (int)pp[u].y, al[3]);
    }
  }
  // fold the BPW block sub-groups of a warp (D = 64), then the warps
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = LPB; o < 32; o <<= 1) {
      ah[i] += __shfl_xor_sync(0xffffffffu, ah[i], o);
      al[i] += __shfl_xor_sync(0xffffffffu, al[i], o);
    }
  }
  if (bsub == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      red[(warp * D + lb * 4 + i) * 2] = ah[i];
      red[(warp * D + lb * 4 + i) * 2 + 1] = al[i];
    }
  }
  __syncthreads();
  const float oscale = pm * (1.f / 127.f) / sum;
  for (int d = tid; d < D; d += 128) {
    int hi = 0, lo = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) hi += red[(w * D + d) * 2], lo += red[(w * D + d) * 2 + 1];
    orow[d] = __float2bfloat16_rn(((float)hi + (float)lo * (1.f / 254.f)) * oscale);
  }
}

size_t lm_attention_q8_smem(int D, int ctxp) { return (size_t)(ctxp + ctxp / 2 + 4 * D * 2) * sizeof(float); }

void launch_lm_attention_q8(const __nv_bfloat16* q, const RingState& rs, const uint8_t* mask, int B, int H, int D, int context,
                            const KvRingQ8& r, __nv_bfloat16* out, cudaStream_t st) {
  const size_t smem = lm_attention_q8_smem(D, r.ctxp);
  DSM_CHECK(smem <= 48 * 1024, DSM_ERR_UNSUPPORTED, "lm_attention_q8: context %d needs %zu bytes of shared memory", context, smem);
#define Q8_ARGS q, rs, mask, H, context, r.ctxp, r.k, r.v, r.k_scale, r.v_scale, r.v_bf16, out
  if (D == 128 && !r.v_bf16) launch_k(lm_attention_q8_kernel<128, 8, true>, B * H, 128, smem, st, Q8_ARGS);
  else if (D == 128) launch_k(lm_attention_q8_kernel<128, 8, false>, B * H, 128, smem, st, Q8_ARGS);
  else if (D == 64 && !r.v_bf16) launch_k(lm_attention_q8_kernel<64, 8, true>, B * H, 128, smem, st, Q8_ARGS);
  else if (D == 64) launch_k(lm_attention_q8_kernel<64, 8, false>, B * H, 128, smem, st, Q8_ARGS);
#undef Q8_ARGS
  else
    fail(DSM_ERR_UNSUPPORTED, "lm_attention_q8: head_dim %d (64 or 128)", D);
  check_launch("lm_attention_q8");
}

// ================================================================================================
// SiLU gate (batched_transformer.rs:170-176): contiguous halves, gate first.
// ================================================================================================
__global__ void silu_gate_kernel(const float* __restrict__ h, int Hd, __nv_bfloat16* __restrict__ g, long long total4) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int per_row = Hd / 4;
  const long long b = i / per_row;
  const int c = (int)(i - b * per_row) * 4;
  const float4 a = *reinterpret_cast<const float4*>(h + b * 2 * Hd + c);
  const float4 v = *reinterpret_cast<const float4*>(h + b * 2 * Hd + Hd + c);
  __nv_bfloat162 lo = __floats2bfloat162_rn(silu(a.x) * v.x, silu(a.y) * v.y);
  __nv_bfloat162 hi = __floats2bfloat162_rn(silu(a.z) * v.z, silu(a.w) * v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&lo);
  u.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(g + b * Hd + c) = u;
}
void launch_silu_gate(const float* h, int B, int Hd, __nv_bfloat16* g, cudaStream_t st) {
  DSM_CHECK(Hd % 4 == 0, DSM_ERR_UNSUPPORTED, "silu_gate: hidden=%d", Hd);
  const long long total4 = (long long)B * Hd / 4;
  launch_k(silu_gate_kernel, (unsigned)ceil_div(total4, 256LL), 256, 0, st, h, Hd, g, total4);
  check_launch("silu_gate");
}

// ================================================================================================
// greedy argmax over the text logits (asr.rs:208-210), first index on ties.
// ================================================================================================
// The text head may arrive as split-K partial planes [S][B][V]: they are added here in split order and the finished
// logits are written out (they are an output of LmModel::forward).
__global__ void __launch_bounds__(256) argmax_kernel(float* __restrict__ logits, const float* __restrict__ partials, int S,
                                                      long long part_stride, int V, uint32_t* __restrict__ tokens) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  float* row = logits + (long long)b * V;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  if (S > 0 && V % 4 == 0) {
    for (int j = tid * 4; j < V; j += blockDim.x * 4) {
      float4 q[kMaxPartials];
#pragma unroll
      for (int sp = 0; sp < kMaxPartials; ++sp)
        if (sp < S) q[sp] = __ldcg(reinterpret_cast<const float4*>(partials + sp * part_stride + (long long)b * V + j));
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int sp = 0; sp < kMaxPartials; ++sp)
        if (sp < S) a.x += q[sp].x, a.y += q[sp].y, a.z += q[sp].z, a.w += q[sp].w;
      *reinterpret_cast<float4*>(row + j) = a;
      if (a.x > best) best = a.x, bi = j;
      if (a.y > best) best = a.y, bi = j + 1;
      if (a.z > best) best = a.z, bi = j + 2;
      if (a.w > best) best = a.w, bi = j + 3;
    }
  } else {
    for (int j = tid; j < V; j += blockDim.x) {
      float v;
      if (S > 0) {
        v = 0.f;
        for (int sp = 0; sp < S; ++sp) v += __ldcg(partials + sp * part_stride + (long long)b * V + j);
        row[j] = v;
      } else {
        v = row[j];
      }
      if (v > best) best = v, bi = j;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
  }
  if ((tid & 31) == 0) s_val[tid >> 5] = best, s_idx[tid >> 5] = bi;
  __syncthreads();
  if (tid == 0) {
    best = s_val[0], bi = s_idx[0];
    for (int w = 1; w < 8; ++w)
      if (s_val[w] > best || (s_val[w] == best && s_idx[w] < bi)) best = s_val[w], bi = s_idx[w];
    tokens[b] = (bi < 0 || bi >= V) ? 0u : (uint32_t)bi;
  }
}
void launch_argmax(float* logits, const PartialSums& ps, int B, int V, uint32_t* tokens, cudaStream_t st) {
  launch_k(argmax_kernel, B, 256, 0, st, logits, ps.planes, ps.count, ps.stride, V, tokens);
  check_launch("argmax");
}

// ================================================================================================
// extra (VAD) heads: nh x dim tiny GEMV + softmax, class 0 (lm.rs:948-955, asr.rs:195-206).
// One CTA per stream, one warp per (head, class) output.
// ================================================================================================
__global__ void __launch_bounds__(256) extra_heads_kernel(const __nv_bfloat16* __restrict__ ys,
                                                           const __nv_bfloat16* __restrict__ w, int nh, int dim, int B,
                                                           int D, float* __restrict__ prs) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_logit[];  // [nh*dim]
  const int b = blockIdx.x, warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const __nv_bfloat16* y = ys + (long long)b * D;
  for (int o = warp; o < nh * dim; o += blockDim.x / 32) {
    const __nv_bfloat16* wr = w + (long long)o * D;
    float acc = 0.f;
    for (int c = lane * 8; c < D; c += 256) {
      const uint4 a = *reinterpret_cast<const uint4*>(y + c);
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(wr + c));
      acc = fmaf(bf16_lo(a.x), bf16_lo(q.x), acc), acc = fmaf(bf16_hi(a.x), bf16_hi(q.x), acc);
      acc = fmaf(bf16_lo(a.y), bf16_lo(q.y), acc), acc = fmaf(bf16_hi(a.y), bf16_hi(q.y), acc);
      acc = fmaf(bf16_lo(a.z), bf16_lo(q.z), acc), acc = fmaf(bf16_hi(a.z), bf16_hi(q.z), acc);
      acc = fmaf(bf16_lo(a.w), bf16_lo(q.w), acc), acc = fmaf(bf16_hi(a.w), bf16_hi(q.w), acc);
    }
    acc = warp_sum(acc);
    // the reference's head output is a bf16 tensor, widened to f32 before the softmax (asr.rs:199-200)
    if (lane == 0) s_logit[o] = __bfloat162float(__float2bfloat16_rn(acc));
  }
  __syncthreads();
  if (threadIdx.x < nh) {
    const float* l = s_logit + threadIdx.x * dim;
    float mx = l[0];
    for (int i = 1; i < dim; ++i) mx = fmaxf(mx, l[i]);
    float sum = 0.f;
    for (int i = 0; i < dim; ++i) sum += expf(l[i] - mx);
    prs[(long long)threadIdx.x * B + b] = expf(l[0] - mx) / sum;
  }
}
void launch_extra_heads(const __nv_bfloat16* ys, const __nv_bfloat16* w, int nh, int dim, int B, int D, float* prs,
                        cudaStream_t st) {
  DSM_CHECK(D % 256 == 0 && nh <= 256, DSM_ERR_UNSUPPORTED, "extra_heads: D=%d nh=%d", D, nh);
  launch_k(extra_heads_kernel, B, 256, nh * dim * sizeof(float), st, ys, w, nh, dim, B, D, prs);
  check_launch("extra_heads");
}

// ================================================================================================
// ASR token feed / commit (asr.rs:133-145, 168-183, 218-226)
// ================================================================================================
__global__ void asr_prepare_kernel(const uint32_t* __restrict__ next_cb, const uint8_t* __restrict__ first_step,
                                   const uint32_t* __restrict__ text_tok, uint32_t pad, uint32_t start, int B, int n_cb,
                                   uint32_t* __restrict__ lm_audio_in, uint32_t* __restrict__ lm_text_in) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n_cb) return;
  const int b = i / n_cb;
  const bool first = first_step[b] != 0;
  lm_audio_in[i] = first ? pad : next_cb[i];
  if (i == b * n_cb) lm_text_in[b] = first ? start : text_tok[b];
}
void launch_asr_prepare(const uint32_t* next_cb, const uint8_t* first_step, const uint32_t* text_tok, uint32_t pad,
                        uint32_t start, int B, int n_cb, uint32_t* lm_audio_in, uint32_t* lm_text_in, cudaStream_t st) {
  launch_k(asr_prepare_kernel, ceil_div(B * n_cb, 256), 256, 0, st, next_cb, first_step, text_tok, pad, start, B, n_cb,
                                                             lm_audio_in, lm_text_in);
  check_launch("asr_prepare");
}

__global__ void asr_commit_kernel(const uint32_t* __restrict__ codes, const uint8_t* __restrict__ mask,
                                  const uint32_t* __restrict__ sampled, int B, int n_cb, uint32_t* __restrict__ next_cb,
                                  uint32_t* __restrict__ text_tok, uint8_t* __restrict__ first_step) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n_cb) return;
  const int b = i / n_cb;
  if (mask && !mask[b]) return;
  next_cb[i] = codes[i];
  if (i == b * n_cb) {
    text_tok[b] = sampled[b];
    first_step[b] = 0;
  }
}
void launch_asr_commit(const uint32_t* codes, const uint8_t* mask, const uint32_t* sampled, int B, int n_cb,
                       uint32_t* next_cb, uint32_t* text_tok, uint8_t* first_step, cudaStream_t st) {
  launch_k(asr_commit_kernel, ceil_div(B * n_cb, 256), 256, 0, st, codes, mask, sampled, B, n_cb, next_cb, text_tok, first_step);
  check_launch("asr_commit");
}

}  // namespace dsm
