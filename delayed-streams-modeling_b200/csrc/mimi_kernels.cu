// Hand-written kernels of the Mimi codec path that are not GEMM-shaped: ring bookkeeping, the C_in=1
// and C_out=1 SEANet edge convs, LayerNorm, RoPE + ring scatter, ring attention (f32), RVQ argmin /
// residual update, RVQ dequant gather, depthwise upsample, conv state commit.
#include "kernels.cuh"

namespace dsm {

// ================================================================================================
// ring bookkeeping -- kv_cache.rs:119-237 on the device, any t, no host sync
// ================================================================================================
__global__ void ring_advance_kernel(RingState rs, const uint8_t* __restrict__ mask, int B, int t, int context) {
  pdl_launch_dependents();
  pdl_wait();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const bool active = mask ? mask[b] != 0 : true;
  uint32_t i0 = rs.indices[b], p0 = rs.positions[b];
  rs.cur_index[b] = i0;
  rs.cur_pos[b] = p0;
  if (active) {
    i0 = (i0 + t) % context;
    p0 += t;
    rs.indices[b] = i0;
    rs.positions[b] = p0;
  }
  // batched_transformer.rs:440-447 reads positions after indices_and_mask advanced them
  rs.rope_pos[b] = p0;
}

void launch_ring_advance(const RingState& rs, const uint8_t* mask, int B, int t, int context, cudaStream_t st) {
  launch_k(ring_advance_kernel, ceil_div(B, 128), 128, 0, st, rs, mask, B, t, context);
  check_launch("ring_advance");
}

// What the attention kernels see, written out (known-answer check against kv_cache.rs:338-405).  t == 1 uses the
// LM kernels' fast-path rule, t > 1 the general one of the codec kernel.
__global__ void ring_debug_kernel(RingState rs, const uint8_t* __restrict__ mask, int t, int context,
                                  uint32_t* __restrict__ indices, float* __restrict__ amask) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const bool active = mask ? mask[b] != 0 : true;
  const int i0 = (int)rs.cur_index[b];
  const uint32_t p0 = rs.cur_pos[b];
  for (int s = threadIdx.x; s < t; s += blockDim.x)
    indices[b * t + s] = active ? (uint32_t)ring_write_slot(i0, s, context) : (uint32_t)i0;
  for (int i = threadIdx.x; i < t * context; i += blockDim.x) {
    const int s = i / context, j = i - s * context;
    bool vis;
    if (t == 1) vis = !active || j < ring_visible_count(p0, context);
    else vis = ring_slot_visible(j, i0, (long long)p0, t, context, active, s);
    amask[((long long)b * t + s) * context + j] = vis ? 0.f : -INFINITY;
  }
}
void launch_ring_debug(const RingState& rs, const uint8_t* mask, int B, int t, int context, uint32_t* indices, float* amask,
                       cudaStream_t st) {
  launch_k(ring_debug_kernel, B, 128, 0, st, rs, mask, t, context, indices, amask);
  check_launch("ring_debug");
}

// ================================================================================================
// SEANet first conv: C_in = 1, k taps, stride 1 (seanet.rs:177-190).  HBM-bound on its output
// ([B, T, C] f32 written twice: raw for the residual path, ELU'd for the next conv).
// One thread produces 4 consecutive channels of one time step -> float4 coalesced stores.
// ================================================================================================
constexpr int IC_R = 16;   // time steps per thread: 256-step tiles amortise the weight staging of a CTA
template <int KMAX>
__global__ void __launch_bounds__(256) init_conv_kernel(const float* __restrict__ pcm, const float* __restrict__ state,
                                                         int S, const float* __restrict__ w,
                                                         const float* __restrict__ bias, int T, int C, int k,
                                                         float* __restrict__ out_raw, RowMap raw_map,
                                                         float* __restrict__ out_act, RowMap act_map) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  float* sw = sm;                 // [C][k]
  float* sx = sm + C * k;         // [tile_t + k - 1]
  const int b = blockIdx.y;
  const int c4n = C / 4;
  const int tile_t = blockDim.x / c4n * IC_R;   // each thread handles IC_R time steps
  const int t0 = blockIdx.x * tile_t;
  for (int i = threadIdx.x; i < C * k; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < tile_t + k - 1; i += blockDim.x) {
    int ti = t0 + i;  // index into [state | pcm]
    float v = 0.f;
    if (ti < S) v = state[(long long)b * S + ti];
    else if (ti - S < T) v = pcm[(long long)b * T + ti - S];
    sx[i] = v;
  }
  __syncthreads();
  const int c4 = threadIdx.x % c4n;
  const int tt = threadIdx.x / c4n;
  float wreg[4][KMAX];
  float breg[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    breg[j] = bias ? bias[c4 * 4 + j] : 0.f;
#pragma unroll
    for (int q = 0; q < KMAX; ++q) wreg[j][q] = q < k ? sw[(c4 * 4 + j) * k + q] : 0.f;
  }
#pragma unroll 4
  for (int r = 0; r < IC_R; ++r) {
    const int tl = tt + r * (tile_t / IC_R);
    const int t = t0 + tl;
    if (t >= T) continue;
    float acc[4] = {breg[0], breg[1], breg[2], breg[3]};
#pragma unroll
    for (int q = 0; q < KMAX; ++q) {
      if (q < k) {
        float x = sx[tl + q];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(wreg[j][q], x, acc[j]);
      }
    }
    const int m = b * T + t;
    if (out_raw)
      *reinterpret_cast<float4*>(out_raw + raw_map.off(m) + c4 * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    if (out_act)
      *reinterpret_cast<float4*>(out_act + act_map.off(m) + c4 * 4) =
          make_float4(elu1(acc[0]), elu1(acc[1]), elu1(acc[2]), elu1(acc[3]));
  }
}

void launch_init_conv(const float* pcm, const float* state, int S, const float* w, const float* bias, int B, int T,
                      int C, int k, float* out_raw, RowMap raw_map, float* out_act, RowMap act_map, cudaStream_t st) {
  DSM_CHECK(k <= 8 && C % 4 == 0 && 256 % (C / 4) == 0, DSM_ERR_UNSUPPORTED, "init_conv: k=%d C=%d unsupported", k, C);
  const int tile_t = 256 / (C / 4) * IC_R;
  dim3 grid(ceil_div(T, tile_t), B);
  size_t smem = (size_t)(C * k + tile_t + k - 1) * sizeof(float);
  launch_k(init_conv_kernel<8>, grid, 256, smem, st, pcm, state, S, w, bias, T, C, k, out_raw, raw_map, out_act, act_map);
  check_launch("init_conv");
}

// ================================================================================================
// SEANet decoder last conv: C_out = 1 (seanet.rs:381-398).  Reads the ELU'd [B, S+T, C] buffer once.
// One warp per output sample: lanes stride over the contiguous k*C window.
// ================================================================================================
__global__ void __launch_bounds__(256) final_conv_kernel(const float* __restrict__ x, RowMap x_map,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          int B, int T, int KC, float* __restrict__ pcm) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sw[];
  for (int i = threadIdx.x; i < KC; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long long m = (long long)blockIdx.x * (blockDim.x / 32) + warp;
  if (m >= (long long)B * T) return;
  const float* row = x + x_map.off((int)m);
  float acc = 0.f;
  for (int i = lane * 4; i < KC; i += 128) {
    float4 v = *reinterpret_cast<const float4*>(row + i);
    acc = fmaf(v.x, sw[i], acc);
    acc = fmaf(v.y, sw[i + 1], acc);
    acc = fmaf(v.z, sw[i + 2], acc);
    acc = fmaf(v.w, sw[i + 3], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) pcm[m] = acc + (bias ? bias[0] : 0.f);
}

void launch_final_conv(const float* x, RowMap x_map, const float* w, const float* bias, int B, int T, int C, int k,
                       float* pcm, cudaStream_t st) {
  const int KC = k * C;
  DSM_CHECK(KC % 4 == 0, DSM_ERR_UNSUPPORTED, "final_conv: k*C=%d", KC);
  launch_k(final_conv_kernel, ceil_div(B * T, 8), 256, KC * sizeof(float), st, x, x_map, w, bias, B, T, KC, pcm);
  check_launch("final_conv");
}

// ================================================================================================
// conv state commit / replicate pad / per-slot zero
// ================================================================================================
// One CTA per (entry, stream).  Values are staged through registers so a state longer than the step
// (source and destination ranges overlap: the state shifts by T rows) is still copied correctly.
__global__ void __launch_bounds__(256) commit_state_kernel(const CommitEntry* __restrict__ entries,
                                                            const uint8_t* __restrict__ mask, int* first_step_flag) {
  pdl_launch_dependents();
  pdl_wait();
  const CommitEntry e = entries[blockIdx.x];
  const int b = blockIdx.y;
  const bool active = mask ? mask[b] != 0 : true;
  float* dst = e.dst + (long long)b * e.dst_stride_b;
  if (active) {
    const float* src = e.src + (long long)b * e.src_stride_b;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int idx = threadIdx.x + i * 256;
      v[i] = idx < e.count ? src[idx] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int idx = threadIdx.x + i * 256;
      if (idx < e.count) dst[idx] = v[i];
    }
  } else if (e.zero_if_inactive_first && first_step_flag && *first_step_flag) {
    // conv.rs:351-355: a stream masked out on the module's first step starts from zero state
    for (int i = threadIdx.x; i < e.count; i += blockDim.x) dst[i] = 0.f;
  }
}
__global__ void clear_flag_kernel(int* flag) {
  pdl_launch_dependents();
  pdl_wait();
  *flag = 0;
}

void launch_commit_state(const CommitEntry* entries_dev, int n_entries, int max_count, const uint8_t* mask,
                         int* first_step_flag, int B, cudaStream_t st) {
  DSM_CHECK(max_count <= 4096, DSM_ERR_UNSUPPORTED, "commit_state: %d floats per entry (max 4096)", max_count);
  launch_k(commit_state_kernel, dim3(n_entries, B), 256, 0, st, entries_dev, mask, first_step_flag);
  check_launch("commit_state");
  if (first_step_flag) {
    launch_k(clear_flag_kernel, 1, 1, 0, st, first_step_flag);
    check_launch("clear_flag");
  }
}

__global__ void replicate_pad_kernel(float* x, long long stride_b, int S, int C, const int* first_step_flag) {
  pdl_launch_dependents();
  pdl_wait();
  if (!*first_step_flag) return;
  float* base = x + (long long)blockIdx.x * stride_b;
  for (int i = threadIdx.x; i < S * C; i += blockDim.x) base[i] = base[S * C + (i % C)];
}
void launch_replicate_pad(float* x, long long stride_b, int S, int C, const int* first_step_flag, int B,
                          cudaStream_t st) {
  launch_k(replicate_pad_kernel, B, 256, 0, st, x, stride_b, S, C, first_step_flag);
  check_launch("replicate_pad");
}

__global__ void zero_rows_kernel(const CommitEntry* __restrict__ entries, int b0) {
  pdl_launch_dependents();
  pdl_wait();
  const CommitEntry e = entries[blockIdx.x];
  float* dst = e.dst + (long long)(b0 + blockIdx.y) * e.dst_stride_b;
  for (int i = threadIdx.x; i < e.count; i += blockDim.x) dst[i] = 0.f;
}
void launch_zero_rows(const CommitEntry* entries_dev, int n_entries, int b0, int nb, cudaStream_t st) {
  launch_k(zero_rows_kernel, dim3(n_entries, nb), 256, 0, st, entries_dev, b0);
  check_launch("zero_rows");
}

// ================================================================================================
// LayerNorm (candle_nn::LayerNorm via batched_transformer.rs:200-222): one warp per row, two-pass in
// registers, biased variance.
// ================================================================================================
template <int D>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, RowMap x_map,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         float eps, int M, float* __restrict__ out,
                                                         __nv_bfloat16* __restrict__ planes, long long plane_stride) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int V = D / 128;  // float4 per lane
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int m = blockIdx.x * (blockDim.x / 32) + warp;
  if (m >= M) return;
  const float* row = x + x_map.off(m);
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = *reinterpret_cast<const float4*>(row + (i * 32 + lane) * 4);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
  float* orow = out + (long long)m * D;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = (i * 32 + lane) * 4;
    float4 g = *reinterpret_cast<const float4*>(w + c);
    float4 bb = *reinterpret_cast<const float4*>(bias + c);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + bb.x;
    o.y = (v[i].y - mean) * rstd * g.y + bb.y;
    o.z = (v[i].z - mean) * rstd * g.z + bb.z;
    o.w = (v[i].w - mean) * rstd * g.w + bb.w;
    *reinterpret_cast<float4*>(orow + c) = o;
    if (planes) store_planes4(planes, plane_stride, (long long)m * D + c, o);
  }
}

void launch_layernorm(const float* x, RowMap x_map, const float* w, const float* b, float eps, int M, int D,
                      float* out, cudaStream_t st, __nv_bfloat16* planes, long long plane_stride) {
  DSM_CHECK(D == 512, DSM_ERR_UNSUPPORTED, "layernorm: D=%d (only 512)", D);
  launch_k(layernorm_kernel<512>, ceil_div(M, 8), 256, 0, st, x, x_map, w, b, eps, M, out, planes, plane_stride);
  check_launch("layernorm");
}

// ================================================================================================
// RoPE (interleaved pairs, transformer.rs:366-403) + scatter of k, v into the ring (kv_cache.rs:28-42)
// inv_freq [D/2] is a host-computed table (libm powf, like the reference's f32::powf) so the angles match it.
// ================================================================================================
__global__ void __launch_bounds__(256) rope_scatter_f32_kernel(float* __restrict__ qkv, RingState rs,
                                                                const uint8_t* __restrict__ mask, int t, int H, int D,
                                                                int context, const float* __restrict__ inv_freq,
                                                                float* __restrict__ k_ring,
                                                                float* __restrict__ v_ring) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x;
  const int b = m / t, s = m - b * t;
  const bool active = mask ? mask[b] != 0 : true;
  const uint32_t i0 = rs.cur_index[b];
  // kv_cache.rs:192-196: an inactive stream's indices repeat its (un-advanced) write slot
  const int slot = active ? (int)((i0 + s) % context) : (int)i0;
  const float pos = (float)(rs.rope_pos[b] + (uint32_t)s);
  const int HD = H * D, half = D / 2;
  float* row = qkv + (long long)m * 3 * HD;
  for (int i = threadIdx.x; i < H * half; i += blockDim.x) {
    const int h = i / half, p = i - h * half;
    float sn, cs;
    sincosf(pos * inv_freq[p], &sn, &cs);
    float2 q = *reinterpret_cast<float2*>(row + h * D + 2 * p);
    float2 k = *reinterpret_cast<float2*>(row + HD + h * D + 2 * p);
    float2 v = *reinterpret_cast<float2*>(row + 2 * HD + h * D + 2 * p);
    float2 qr = make_float2(q.x * cs - q.y * sn, q.x * sn + q.y * cs);
    float2 kr = make_float2(k.x * cs - k.y * sn, k.x * sn + k.y * cs);
    *reinterpret_cast<float2*>(row + h * D + 2 * p) = qr;
    const long long o = (((long long)b * H + h) * context + slot) * D + 2 * p;
    *reinterpret_cast<float2*>(k_ring + o) = kr;
    *reinterpret_cast<float2*>(v_ring + o) = v;
  }
}

void launch_rope_scatter_f32(float* qkv, const RingState& rs, const uint8_t* mask, int B, int t, int H, int D,
                             int context, const float* inv_freq, float* k_ring, float* v_ring, cudaStream_t st) {
  launch_k(rope_scatter_f32_kernel, B * t, 256, 0, st, qkv, rs, mask, t, H, D, context, inv_freq, k_ring, v_ring);
  check_launch("rope_scatter_f32");
}

// ================================================================================================
// Codec ring attention, f32, D = 64, t <= 2 queries per stream (batched_transformer.rs:107-113 with the
// mask of kv_cache.rs:192-229 recomputed from (cur_index, cur_pos) instead of materialised).
// One CTA per (stream, head); K and V rows are read once for both queries.
// ================================================================================================
template <int D, int TQ>
__global__ void __launch_bounds__(128) codec_attention_kernel(const float* __restrict__ qkv, RingState rs,
                                                               const uint8_t* __restrict__ mask, int t, int H,
                                                               int context, const float* __restrict__ k_ring,
                                                               const float* __restrict__ v_ring,
                                                               float* __restrict__ out,
                                                               __nv_bfloat16* __restrict__ planes, long long plane_stride) {
  pdl_launch_dependents();
  pdl_wait();
  static_assert(D == 64, "lane layout assumes 16 lanes x float4 per key");
  extern __shared__ float sm[];
  float* sc = sm;                       // [TQ][context]
  float* red = sm + TQ * context;       // [8][TQ][D]
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const bool active = mask ? mask[b] != 0 : true;
  const int i0 = (int)rs.cur_index[b];
  const long long p0 = rs.cur_pos[b];
  const int HD = H * D;
  const float scale = rsqrtf((float)D);

  // this lane's chunk of each query
  const int c4 = lane % 16, sub = lane / 16;
  float4 q[TQ];
#pragma unroll
  for (int s = 0; s < TQ; ++s)
    q[s] = s < t ? *reinterpret_cast<const float4*>(qkv + ((long long)(b * t + s) * 3) * HD + h * D + c4 * 4)
                 : make_float4(0.f, 0.f, 0.f, 0.f);

  const float* kb = k_ring + ((long long)b * H + h) * context * D;
  const float* vb = v_ring + ((long long)b * H + h) * context * D;

  // Scores.  4 warps x 2 keys x UNR keys in flight per iteration; 16 lanes share a key (one float4 each), so
  // every iteration leaves 2*UNR partial dot products per lane that must be summed over 16 lanes.  They are
  // reduced together: at each butterfly step a lane keeps half of its values and hands the other half to its
  // partner (8 shuffles per iteration instead of 32; the pairing order 8,4,2,1 and therefore the rounding are
  // those of a plain xor-shuffle tree).  The visibility mask is applied later, once per score.
  constexpr int UNR = 8;     // keys in flight per lane group: 8 x 16 B per lane, ~115 KB per SM at 7 CTAs
  static_assert(TQ == 2 && UNR % 4 == 0, "the multi-value reduction below takes 4 keys x 2 queries at a time");
  const bool h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
  const int my_u = (h8 ? 2 : 0) + (h4 ? 1 : 0), my_s = h2 ? 1 : 0;   // the (key, query) this lane ends up with
  for (int j0 = warp * 2 * UNR; j0 < context; j0 += 8 * UNR) {
    float4 kv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + 2 * u + sub;
      kv[u] = j < context ? __ldg(reinterpret_cast<const float4*>(kb + (long long)j * D + c4 * 4))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int ub = 0; ub < UNR; ub += 4) {
    float v8[8];   // index u * 2 + s
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int s = 0; s < TQ; ++s)
        v8[u * 2 + s] = q[s].x * kv[ub + u].x + q[s].y * kv[ub + u].y + q[s].z * kv[ub + u].z + q[s].w * kv[ub + u].w;
    float v4[4], v2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float recv = __shfl_xor_sync(0xffffffffu, h8 ? v8[i] : v8[i + 4], 8);
      v4[i] = (h8 ? v8[i + 4] : v8[i]) + recv;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float recv = __shfl_xor_sync(0xffffffffu, h4 ? v4[i] : v4[i + 2], 4);
      v2[i] = (h4 ? v4[i + 2] : v4[i]) + recv;
    }
    float v1 = (h2 ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, h2 ? v2[0] : v2[1], 2);
    v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
    const int j = j0 + 2 * (ub + my_u) + sub;
    if ((lane & 1) == 0 && j < context && my_s < t) sc[my_s * context + j] = v1 * scale;
    }
  }
  __syncthreads();

  // mask + softmax: warp s normalises query s.  Slot positions as kv_cache.rs:199-214 reconstructs them,
  // overridden for the slots written this step.
  if (warp < t) {
    float* row = sc + warp * context;
    auto visible = [&](int j) -> bool { return ring_slot_visible(j, i0, p0, t, context, active, warp); };
    float mx = -INFINITY;
    for (int j = lane; j < context; j += 32) {
      const float x = visible(j) ? row[j] : -INFINITY;
      row[j] = x;
      mx = fmaxf(mx, x);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < context; j += 32) {
      float e = expf(row[j] - mx);
      row[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < context; j += 32) row[j] *= inv;
  }
  __syncthreads();

  // P.V: thread = (key group g of 8, float4 chunk d4 of 16)
  const int d4 = tid % 16, g = tid / 16;
  float4 acc[TQ];
#pragma unroll
  for (int s = 0; s < TQ; ++s) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j0 = g; j0 < context; j0 += 8 * UNR) {
    float4 vv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + 8 * u;
      vv[u] = j < context ? __ldg(reinterpret_cast<const float4*>(vb + (long long)j * D + d4 * 4))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + 8 * u;
      if (j >= context) break;
#pragma unroll
      for (int s = 0; s < TQ; ++s) {
        const float p = sc[s * context + j];
        acc[s].x = fmaf(p, vv[u].x, acc[s].x);
        acc[s].y = fmaf(p, vv[u].y, acc[s].y);
        acc[s].z = fmaf(p, vv[u].z, acc[s].z);
        acc[s].w = fmaf(p, vv[u].w, acc[s].w);
      }
    }
  }
#pragma unroll
  for (int s = 0; s < TQ; ++s) *reinterpret_cast<float4*>(red + ((g * TQ + s) * D) + d4 * 4) = acc[s];
  __syncthreads();
  for (int i = tid; i < t * D; i += blockDim.x) {
    const int s = i / D, d = i - s * D;
    float v = 0.f;
#pragma unroll
    for (int gg = 0; gg < 8; ++gg) v += red[(gg * TQ + s) * D + d];
    const long long o = (long long)(b * t + s) * HD + h * D + d;
    out[o] = v;
    if (planes) {
      uint16_t ph, pm, pl;
      split3_bf16(v, ph, pm, pl);
      reinterpret_cast<uint16_t*>(planes)[o] = ph;
      reinterpret_cast<uint16_t*>(planes)[plane_stride + o] = pm;
      reinterpret_cast<uint16_t*>(planes)[2 * plane_stride + o] = pl;
    }
  }
}

void launch_codec_attention(const float* qkv, const RingState& rs, const uint8_t* mask, int B, int t, int H, int D,
                            int context, const float* k_ring, const float* v_ring, float* out, cudaStream_t st,
                            __nv_bfloat16* planes, long long plane_stride) {
  DSM_CHECK(D == 64 && t <= 2, DSM_ERR_UNSUPPORTED, "codec_attention: D=%d t=%d", D, t);
  size_t smem = (size_t)(2 * context + 8 * 2 * D) * sizeof(float);
  launch_k(codec_attention_kernel<64, 2>, B * H, 128, smem, st, qkv, rs, mask, t, H, context, k_ring, v_ring, out, planes,
                                                          plane_stride);
  check_launch("codec_attention");
}

// ================================================================================================
// RVQ: first-index argmin over the (c2 - dot) scores, residual update (quantization.rs:122-131,
// 182-194, 219-229).  One CTA per stream.
// ================================================================================================
__global__ void __launch_bounds__(256) rvq_select_kernel(const float* __restrict__ scores, int bins,
                                                          const float* __restrict__ embedding, int qdim,
                                                          float* __restrict__ residual, long long res_stride,
                                                          uint32_t* __restrict__ codes, int n_q, int q,
                                                          __nv_bfloat16* __restrict__ planes, long long plane_stride) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* row = scores + (long long)b * bins;
  float best = INFINITY;
  int bi = 0x7fffffff;
  for (int j = tid; j < bins; j += blockDim.x) {
    float v = row[j];
    if (v < best) best = v, bi = j;   // ascending j per thread: strict < keeps the first
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov < best || (ov == best && oi < bi)) best = ov, bi = oi;
  }
  if ((tid & 31) == 0) s_val[tid >> 5] = best, s_idx[tid >> 5] = bi;
  __syncthreads();
  best = s_val[0], bi = s_idx[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) {
    float ov = s_val[w];
    int oi = s_idx[w];
    if (ov < best || (ov == best && oi < bi)) best = ov, bi = oi;
  }
  if (bi < 0 || bi >= bins) bi = 0;  // all-NaN row: candle's argmin would also return 0
  if (tid == 0) codes[(long long)b * n_q + q] = (uint32_t)bi;
  const float* e = embedding + (long long)bi * qdim;
  float* r = residual + (long long)b * res_stride;
  for (int d = tid; d < qdim; d += blockDim.x) {
    const float v = r[d] - e[d];
    r[d] = v;
    if (planes) {
      // the next codebook's score GEMM reads the residual as three bf16 planes (same row layout as `residual`)
      const long long o = (long long)b * res_stride + d;
      uint16_t ph, pm, pl;
      split3_bf16(v, ph, pm, pl);
      reinterpret_cast<uint16_t*>(planes)[o] = ph;
      reinterpret_cast<uint16_t*>(planes)[plane_stride + o] = pm;
      reinterpret_cast<uint16_t*>(planes)[2 * plane_stride + o] = pl;
    }
  }
}

void launch_rvq_select(const float* scores, int B, int bins, const float* embedding, int qdim, float* residual,
                       long long res_stride, uint32_t* codes, int n_q, int q, cudaStream_t st, __nv_bfloat16* planes,
                       long long plane_stride) {
  launch_k(rvq_select_kernel, B, 256, 0, st, scores, bins, embedding, qdim, residual, res_stride, codes, n_q, q, planes,
                                       plane_stride);
  check_launch("rvq_select");
}

__global__ void __launch_bounds__(256) rvq_gather_sum_kernel(const uint32_t* __restrict__ codes, int n_q, int q0,
                                                              int q1, const float* __restrict__ embeddings, int bins,
                                                              int qdim, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < qdim; d += blockDim.x) {
    float acc = 0.f;
    for (int q = q0; q < q1; ++q) {
      uint32_t c = codes[(long long)b * n_q + q];
      if (c >= (uint32_t)bins) c = bins - 1;
      // quantization.rs:241-246 accumulates layer by layer in this order
      acc += embeddings[((long long)q * bins + c) * qdim + d];
    }
    out[(long long)b * qdim + d] = acc;
  }
}
void launch_rvq_gather_sum(const uint32_t* codes, int n_q, int q0, int q1, const float* embeddings, int bins, int qdim,
                           int B, float* out, cudaStream_t st) {
  launch_k(rvq_gather_sum_kernel, B, 256, 0, st, codes, n_q, q0, q1, embeddings, bins, qdim, out);
  check_launch("rvq_gather_sum");
}

// ================================================================================================
// Depthwise conv-transpose upsample, k = 2s, stride s (conv.rs:558-606; the reference expands it to a
// dense 512x512x4 weight at load, conv.rs:144-150 -- here it stays depthwise).
// out[b, r, c] = x[b, c] * w[c, r] + prev[b, c] * w[c, r + s]
// ================================================================================================
__global__ void upsample_dw_kernel(const float* __restrict__ x, const float* __restrict__ state,
                                   const float* __restrict__ w, int C, int s, float* __restrict__ out, RowMap out_map) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float xv = x[(long long)b * C + c], pv = state[(long long)b * C + c];
    for (int r = 0; r < s; ++r)
      out[out_map.off(b * s + r) + c] = xv * w[c * 2 * s + r] + pv * w[c * 2 * s + r + s];
  }
}
void launch_upsample_dw(const float* x, const float* state, const float* w, int B, int C, int s, float* out,
                        RowMap out_map, cudaStream_t st) {
  launch_k(upsample_dw_kernel, B, 256, 0, st, x, state, w, C, s, out, out_map);
  check_launch("upsample_dw");
}

}  // namespace dsm
