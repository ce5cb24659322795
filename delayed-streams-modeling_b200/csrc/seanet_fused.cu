// SEANet encoder stage 1 as ONE kernel: init conv (1 -> 64, k 7) -> resnet block (ELU, conv k3 64 -> 32, ELU, conv k1
// 32 -> 64, + skip) -> ELU, i.e. encoder.model.0 and encoder.model.1 (seanet.rs:177-245, 140-150, conv.rs:312-370).
//
// As separate kernels these three layers move the [B, 1920, 64] activation through HBM five times (raw + ELU'd copy out
// of the init conv, ELU'd copy into the k3 conv, raw copy into the k1 conv's skip, ELU'd result out) and the [B, 1920, 32]
// hidden twice; on the tensor-core GEMM they are its worst cases -- K = 192 / N = 32 and K = 32: one or three k-blocks
// per tile, all epilogue (profiles/r02j: 27 % and 19 % of their rooflines at 2,000 streams).  Here a CTA takes 128 time
// steps of one stream from PCM to the block's ELU'd output with everything in shared memory: 7.7 KB of PCM in and
// 492 KB out per stream-frame, 8,640 f32 FMAs per time step on the CUDA cores (no operand splitting: exact f32, the
// same accumulation order as the reference's f32 dot products up to association).
//
// Streaming state stays where the layer-by-layer path keeps it: the left context of the k3 conv is the ELU'd tail of
// the previous frame in the head rows of its input buffer X1 (this kernel reads them for the first tile and writes this
// frame's last two rows behind them for commit_state), the init conv's 6 samples live in state0; within a frame the
// two halo rows of a tile are recomputed from PCM (bit-identical to what the neighbouring tile computed).
#include "kernels.cuh"

namespace dsm {

namespace {
constexpr int F1_TT = 128;       // time steps per CTA
constexpr int F1_C = 64;         // block channels
constexpr int F1_H = 32;         // hidden channels
constexpr int F1_K0 = 7;         // init conv taps
constexpr int F1_LD = F1_TT + 4; // row length of the transposed activation tiles (rows stay 16-byte aligned)
constexpr int F1_THREADS = 128;
constexpr int F1_W0LD = 8;       // init conv weights padded to 8 per channel: [w0 .. w6, bias]
__device__ __forceinline__ float elu_sel(float x) { return x > 0.f ? x : expf(fminf(x, 0.f)) - 1.f; }   // no divergent branch
}  // namespace

struct Stage1Params {
  const float* pcm;      // [B, T]
  const float* state0;   // [B, 6] previous samples
  const float* w0b;      // [64][8]   init conv taps 0..6 and the bias
  const float* w1t;      // [192][32]  (k = tap * 64 + channel, transposed for the kernel)
  const float* b1;       // [32]
  const float* w2t;      // [32][64]
  const float* b2;       // [64]
  float* x1;             // k3 conv's input buffer [B][2 + T][64] (ELU'd): head rows read, tail rows written
  long long x1_stride_b;
  float* out_act;        // ELU'd block output rows (b, t) at out_act + b * out_stride_b + t * out_stride_t
  long long out_stride_b, out_stride_t;
  int T, tiles_per_cta;
};

// Shared memory: the two weight matrices, ELU(a0)^T for rows t0-2 .. t0+TT-1 (later overwritten by ELU(hidden)^T), the
// PCM window and the init conv's weights.  The raw a0 the skip connection adds back is recomputed from PCM in the
// epilogue (7 FMAs per element, the same order as the first pass, so the same bits) instead of costing a second tile.
//
// Thread mapping, third version.  ncu on the second one (each thread a small patch of time steps x channels): FMA pipe
// 43 % busy, issue slots 57 %, 2,900 shared-memory wavefronts per warp and tile with 28 % bank conflicts -- a 128-bit
// LDS is served a quarter-warp at a time, so operand loads whose lanes mostly repeat each other's addresses still
// cost four wavefronts.  Here every warp owns one group of output channels (its weight loads are warp-uniform) and
// its lanes own distinct time steps (activation loads carry 512 distinct bytes per instruction).  Measured: conflicts
// 12.6 M -> 3.0 M, wavefronts 2,900 -> 2,400 per warp-tile (a warp-uniform 128-bit load still costs 2-3, not 1), kernel
// time unchanged (350 vs 339 us): at 12 warps per SM the issue slots, not the LSU, are what is left
// (profiles/r02w_*, r02x_*).
__global__ void __launch_bounds__(F1_THREADS, 3) seanet_stage1_kernel(const Stage1Params p) {
  extern __shared__ __align__(16) float f1_smem[];
  float* sw1t = f1_smem;                          // [192][32]
  float* sw2t = sw1t + 3 * F1_C * F1_H;           // [32][64]
  float* sa = sw2t + F1_H * F1_C;                 // [64][F1_LD]   ELU(a0), index r + 2;  then [32][F1_LD] ELU(hidden), index r
  float* sw0 = sa + F1_C * F1_LD;                 // [64][8]
  float* sb = sw0 + F1_C * F1_W0LD;               // b1 [32], b2 [64]
  float* spcm = sb + F1_H + F1_C;                 // [TT + 8]      pcm for t0-8 .. t0+TT-1
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.y;
  for (int i = tid; i < 3 * F1_C * F1_H / 4; i += F1_THREADS)
    reinterpret_cast<float4*>(sw1t)[i] = __ldg(reinterpret_cast<const float4*>(p.w1t) + i);
  for (int i = tid; i < F1_H * F1_C / 4; i += F1_THREADS)
    reinterpret_cast<float4*>(sw2t)[i] = __ldg(reinterpret_cast<const float4*>(p.w2t) + i);
  for (int i = tid; i < F1_C * F1_W0LD / 4; i += F1_THREADS)
    reinterpret_cast<float4*>(sw0)[i] = __ldg(reinterpret_cast<const float4*>(p.w0b) + i);
  if (tid < F1_H) sb[tid] = p.b1[tid];
  if (tid < F1_C) sb[F1_H + tid] = p.b2[tid];
  const float* pcm = p.pcm + (long long)b * p.T;
  const float* st0 = p.state0 + (long long)b * (F1_K0 - 1);
  float* x1 = p.x1 + (long long)b * p.x1_stride_b;
  float* out = p.out_act + (long long)b * p.out_stride_b;
  for (int tile = 0; tile < p.tiles_per_cta; ++tile) {
    const int t0 = (blockIdx.x * p.tiles_per_cta + tile) * F1_TT;
    if (t0 >= p.T) break;
    __syncthreads();                              // the previous tile's readers are done (and the weights are staged)
    // ---- pcm window t0-8 .. t0+TT-1 ([state0 | pcm]) ----
    for (int i = tid; i < F1_TT + 8; i += F1_THREADS) {
      const int g = t0 - 8 + i;
      float v = 0.f;
      if (g >= 0) v = g < p.T ? pcm[g] : 0.f;
      else if (g >= -(F1_K0 - 1)) v = st0[g + (F1_K0 - 1)];
      spcm[i] = v;
    }
    __syncthreads();
    // ---- init conv: a0[r][c] = b0[c] + sum_q w0[c][q] * x[t0 + r - 6 + q]; thread = one row, all channels ----
    {
      const int r = tid, t = t0 + r;
      float x[F1_K0];
#pragma unroll
      for (int q = 0; q < F1_K0; ++q) x[q] = spcm[r + 2 + q];      // pcm index t - 6 + q  ->  window index - (t0 - 8)
      const bool tail = t >= p.T - 2 && t < p.T;                   // next frame's left context (commit_state moves it)
#pragma unroll 4
      for (int c = 0; c < F1_C; ++c) {
        const float4 wa = *reinterpret_cast<const float4*>(sw0 + c * F1_W0LD);
        const float4 wb = *reinterpret_cast<const float4*>(sw0 + c * F1_W0LD + 4);
        float a = wb.w;
        a = fmaf(wa.x, x[0], a), a = fmaf(wa.y, x[1], a), a = fmaf(wa.z, x[2], a), a = fmaf(wa.w, x[3], a);
        a = fmaf(wb.x, x[4], a), a = fmaf(wb.y, x[5], a), a = fmaf(wb.z, x[6], a);
        const float e = elu_sel(a);
        sa[c * F1_LD + r + 2] = e;
        if (tail) x1[(long long)(t + 2) * F1_C + c] = e;
      }
      // the two rows before the tile: the previous frame's ELU'd tail (head of X1) for the first tile, else recomputed
      const int hr = tid / F1_C - 2, hc = tid % F1_C;               // rows -2, -1
      float e;
      if (t0 == 0) e = x1[(long long)(hr + 2) * F1_C + hc];
      else {
        const float4 wa = *reinterpret_cast<const float4*>(sw0 + hc * F1_W0LD);
        const float4 wb = *reinterpret_cast<const float4*>(sw0 + hc * F1_W0LD + 4);
        const float* xs = spcm + hr + 2;
        float a = wb.w;
        a = fmaf(wa.x, xs[0], a), a = fmaf(wa.y, xs[1], a), a = fmaf(wa.z, xs[2], a), a = fmaf(wa.w, xs[3], a);
        a = fmaf(wb.x, xs[4], a), a = fmaf(wb.y, xs[5], a), a = fmaf(wb.z, xs[6], a);
        e = elu_sel(a);
      }
      sa[hc * F1_LD + hr + 2] = e;
    }
    __syncthreads();
    // ---- conv k3, 64 -> 32, on ELU(a0): warp = 8 hidden channels, lane = 4 consecutive time steps ----
    {
      float hacc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) hacc[i][j] = sb[warp * 8 + j];
#pragma unroll 2
      for (int c = 0; c < F1_C; ++c) {
        const float4 a0 = *reinterpret_cast<const float4*>(sa + c * F1_LD + lane * 4);       // rows 4*lane-2 .. 4*lane+1
        const float2 a1 = *reinterpret_cast<const float2*>(sa + c * F1_LD + lane * 4 + 4);   // rows 4*lane+2, +3
        const float a[6] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y};
#pragma unroll
        for (int tap = 0; tap < 3; ++tap) {
          const float4 w0 = *reinterpret_cast<const float4*>(sw1t + (tap * F1_C + c) * F1_H + warp * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(sw1t + (tap * F1_C + c) * F1_H + warp * 8 + 4);
          const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) hacc[i][j] = fmaf(a[i + tap], w[j], hacc[i][j]);
        }
      }
      __syncthreads();                            // every thread is done reading ELU(a0): the tile is reused
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(sa + (warp * 8 + j) * F1_LD + lane * 4) =
            make_float4(elu_sel(hacc[0][j]), elu_sel(hacc[1][j]), elu_sel(hacc[2][j]), elu_sel(hacc[3][j]));
    }
    __syncthreads();
    // ---- conv k1, 32 -> 64, + skip, ELU: warp = 16 channels, lane = time steps lane, lane + 32, lane + 64, lane + 96 ----
    {
      float acc[4][16];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = sb[F1_H + warp * 16 + j];
#pragma unroll 2
      for (int k = 0; k < F1_H; ++k) {
        float h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = sa[k * F1_LD + lane + 32 * i];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float4 w = *reinterpret_cast<const float4*>(sw2t + k * F1_C + warp * 16 + jj * 4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][jj * 4 + 0] = fmaf(h[i], w.x, acc[i][jj * 4 + 0]);
            acc[i][jj * 4 + 1] = fmaf(h[i], w.y, acc[i][jj * 4 + 1]);
            acc[i][jj * 4 + 2] = fmaf(h[i], w.z, acc[i][jj * 4 + 2]);
            acc[i][jj * 4 + 3] = fmaf(h[i], w.w, acc[i][jj * 4 + 3]);
          }
        }
      }
      // skip: raw a0 for the lane's 4 x 16 patch, recomputed from the PCM window
      float x[4][F1_K0];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < F1_K0; ++q) x[i][q] = spcm[lane + 32 * i + 2 + q];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 wa = *reinterpret_cast<const float4*>(sw0 + (warp * 16 + j) * F1_W0LD);
        const float4 wb = *reinterpret_cast<const float4*>(sw0 + (warp * 16 + j) * F1_W0LD + 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float a = wb.w;
          a = fmaf(wa.x, x[i][0], a), a = fmaf(wa.y, x[i][1], a), a = fmaf(wa.z, x[i][2], a), a = fmaf(wa.w, x[i][3], a);
          a = fmaf(wb.x, x[i][4], a), a = fmaf(wb.y, x[i][5], a), a = fmaf(wb.z, x[i][6], a);
          acc[i][j] = elu_sel(acc[i][j] + a);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = t0 + lane + 32 * i;
        if (t >= p.T) continue;
        float* d = out + (long long)t * p.out_stride_t + warp * 16;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          *reinterpret_cast<float4*>(d + jj * 4) = make_float4(acc[i][jj * 4], acc[i][jj * 4 + 1], acc[i][jj * 4 + 2], acc[i][jj * 4 + 3]);
      }
    }
  }
}

size_t seanet_stage1_smem() {
  return (size_t)(3 * F1_C * F1_H + F1_H * F1_C + F1_C * F1_LD + F1_C * F1_W0LD + F1_H + F1_C + F1_TT + 8) * sizeof(float);
}

bool seanet_stage1_supported(int c0, int k0, int c1h, int k1, int k2, int T) {
  return c0 == F1_C && k0 == F1_K0 && c1h == F1_H && k1 == 3 && k2 == 1 && T % F1_TT == 0;
}

void launch_seanet_stage1(const float* pcm, const float* state0, const float* w0b, const float* w1t, const float* b1,
                          const float* w2t, const float* b2, float* x1, long long x1_stride_b, float* out_act, RowMap out_map, int B,
                          int T, cudaStream_t st) {
  Stage1Params p;
  p.pcm = pcm, p.state0 = state0, p.w0b = w0b, p.w1t = w1t, p.b1 = b1, p.w2t = w2t, p.b2 = b2;
  p.x1 = x1, p.x1_stride_b = x1_stride_b, p.out_act = out_act, p.out_stride_b = out_map.stride_b, p.out_stride_t = out_map.stride_t;
  p.T = T;
  const int tiles = T / F1_TT;
  p.tiles_per_cta = tiles % 3 == 0 ? 3 : 1;       // 15 tiles per frame: 5 CTAs per stream, the weights staged once for three tiles
  const size_t smem = seanet_stage1_smem();
  static PerDeviceOnce attr;
  attr.run([&] { DSM_CUDA(cudaFuncSetAttribute(seanet_stage1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); });
  launch_k(seanet_stage1_kernel, dim3(tiles / p.tiles_per_cta, B), F1_THREADS, smem, st, p);
  check_launch("seanet_stage1");
}

}  // namespace dsm
