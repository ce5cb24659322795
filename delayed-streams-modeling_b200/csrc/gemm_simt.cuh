// f32-accumulate SIMT GEMM  C[m, n] = sum_k A[m, k] * W[n, k]  with row-mapped operands.
//
// This is the CUDA-core workhorse of the Mimi codec, which must stay in true f32 to keep RVQ code
// indices bit-identical to the reference's f32 CPU path (TF32/bf16 tiles flip near-tie argmins).
// Both operands are K-contiguous ("x . W^T", the Linear / implicit-GEMM-conv form):
//   * a streaming conv over an NLC buffer [B, S+T_in, C_in] is this GEMM with K = k*C_in and A-row
//     (b, t) starting at b*stride_b + t*(stride*C_in): the im2col window is already contiguous;
//   * a Linear is the same with T = 1.
// A / W may be f32 or bf16 (bf16 is widened while staging into shared memory).
// The epilogue fuses bias, "c2 - dot" negation (RVQ), LayerScale, residual add, GELU, and writes any
// of: raw f32, ELU(f32) (the next conv's activated input), bf16.
#pragma once

#include "common.cuh"

namespace dsm {

struct GemmParams {
  const void* A = nullptr;
  RowMap a_map{1, 0, 0};
  const void* W = nullptr;  // [N, K]
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;   // [N]
  int neg = 0;                   // v = bias - acc   (instead of acc + bias)
  const float* scale = nullptr;  // [N]  v *= scale[n]
  const float* res = nullptr;    // residual, f32, row-mapped
  RowMap r_map{1, 0, 0};
  int act = 0;                   // 1: GELU(erf)
  float* out_raw = nullptr;
  RowMap o_map{1, 0, 0};
  float* out_act = nullptr;      // ELU(v)
  RowMap e_map{1, 0, 0};
  __nv_bfloat16* out_bf16 = nullptr;
  RowMap h_map{1, 0, 0};
  // tcgen05 codec GEMM only (gemm_tc6): operands pre-split into three bf16 planes, x = p0 + p1 + p2 exactly.
  // a_planes: A as planes (plane q at a_planes + q * a_plane_stride, rows addressed by a_map in elements); the
  // f32 A is then not read.  out_planes: the result additionally stored as planes for the next GEMM.
  const __nv_bfloat16* a_planes = nullptr;
  long long a_plane_stride = 0;
  __nv_bfloat16* out_planes = nullptr;
  RowMap p_map{1, 0, 0};
  long long out_plane_stride = 0;
};

template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
}

template <typename TA, typename TW, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) gemm_simt_kernel(const GemmParams p) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int BK = 16;
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int RM = TM / 4, RN = TN / 4;
  constexpr int KC = BK / 4;                       // float4 chunks per tile row
  constexpr int A_PASSES = (BM * KC + NT - 1) / NT;
  constexpr int W_PASSES = (BN * KC + NT - 1) / NT;
  static_assert(TM % 4 == 0 && TN % 4 == 0, "micro-tile is built from 4x4 blocks");

  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const TA* __restrict__ A = static_cast<const TA*>(p.A);
  const TW* __restrict__ W = static_cast<const TW*>(p.W);

  // per-thread staging assignments
  const TA* a_src[A_PASSES];
  int a_row[A_PASSES], a_kc[A_PASSES];
#pragma unroll
  for (int i = 0; i < A_PASSES; ++i) {
    int id = tid + i * NT;
    a_row[i] = id / KC;
    a_kc[i] = id % KC;
    int m = m0 + a_row[i];
    a_src[i] = (id < BM * KC && m < p.M) ? A + p.a_map.off(m) + a_kc[i] * 4 : nullptr;
  }
  const TW* w_src[W_PASSES];
  int w_row[W_PASSES], w_kc[W_PASSES];
#pragma unroll
  for (int i = 0; i < W_PASSES; ++i) {
    int id = tid + i * NT;
    w_row[i] = id / KC;
    w_kc[i] = id % KC;
    int n = n0 + w_row[i];
    w_src[i] = (id < BN * KC && n < p.N) ? W + (long long)n * p.K + w_kc[i] * 4 : nullptr;
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 a_reg[A_PASSES], w_reg[W_PASSES];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_PASSES; ++i)
      a_reg[i] = a_src[i] ? load4<TA>(a_src[i] + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < W_PASSES; ++i)
      w_reg[i] = w_src[i] ? load4<TW>(w_src[i] + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto sstore = [&]() {
#pragma unroll
    for (int i = 0; i < A_PASSES; ++i) {
      if (tid + i * NT < BM * KC) {
        int k = a_kc[i] * 4, r = a_row[i];
        As[k + 0][r] = a_reg[i].x;
        As[k + 1][r] = a_reg[i].y;
        As[k + 2][r] = a_reg[i].z;
        As[k + 3][r] = a_reg[i].w;
      }
    }
#pragma unroll
    for (int i = 0; i < W_PASSES; ++i) {
      if (tid + i * NT < BN * KC) {
        int k = w_kc[i] * 4, r = w_row[i];
        Ws[k + 0][r] = w_reg[i].x;
        Ws[k + 1][r] = w_reg[i].y;
        Ws[k + 2][r] = w_reg[i].z;
        Ws[k + 3][r] = w_reg[i].w;
      }
    }
  };

  gload(0);
  for (int k0 = 0; k0 < p.K; k0 += BK) {
    sstore();
    __syncthreads();
    if (k0 + BK < p.K) gload(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], w[TN];
#pragma unroll
      for (int g = 0; g < RM; ++g) {
        float4 v = *reinterpret_cast<const float4*>(&As[k][g * (BM / RM) + ty * 4]);
        a[g * 4 + 0] = v.x, a[g * 4 + 1] = v.y, a[g * 4 + 2] = v.z, a[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < RN; ++g) {
        float4 v = *reinterpret_cast<const float4*>(&Ws[k][g * (BN / RN) + tx * 4]);
        w[g * 4 + 0] = v.x, w[g * 4 + 1] = v.y, w[g * 4 + 2] = v.z, w[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ----
#pragma unroll
  for (int gi = 0; gi < RM; ++gi) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + gi * (BM / RM) + ty * 4 + i;
      if (m >= p.M) continue;
      const float* res_row = p.res ? p.res + p.r_map.off(m) : nullptr;
      float* raw_row = p.out_raw ? p.out_raw + p.o_map.off(m) : nullptr;
      float* act_row = p.out_act ? p.out_act + p.e_map.off(m) : nullptr;
      __nv_bfloat16* bf_row = p.out_bf16 ? p.out_bf16 + p.h_map.off(m) : nullptr;
#pragma unroll
      for (int gj = 0; gj < RN; ++gj) {
        const int n = n0 + gj * (BN / RN) + tx * 4;
        if (n >= p.N) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float x = acc[gi * 4 + i][gj * 4 + j];
          if (n + j < p.N) {
            if (p.bias) x = p.neg ? p.bias[n + j] - x : x + p.bias[n + j];
            if (p.act == 1) x = gelu_erf(x);
            if (p.scale) x *= p.scale[n + j];
            if (res_row) x += res_row[n + j];
          }
          v[j] = x;
        }
        if (n + 3 < p.N) {
          if (raw_row) *reinterpret_cast<float4*>(raw_row + n) = make_float4(v[0], v[1], v[2], v[3]);
          if (act_row)
            *reinterpret_cast<float4*>(act_row + n) = make_float4(elu1(v[0]), elu1(v[1]), elu1(v[2]), elu1(v[3]));
          if (bf_row) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]);
            __nv_bfloat162 hi = __floats2bfloat162_rn(v[2], v[3]);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&lo);
            u.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(bf_row + n) = u;
          }
        } else {
          for (int j = 0; j < 4 && n + j < p.N; ++j) {
            if (raw_row) raw_row[n + j] = v[j];
            if (act_row) act_row[n + j] = elu1(v[j]);
            if (bf_row) bf_row[n + j] = __float2bfloat16_rn(v[j]);
          }
        }
      }
    }
  }
}

// Picks a tile shape from (M, N) and launches.  K must be a multiple of 16; all row strides and K
// must keep 16-byte alignment of the 4-element staging loads (true for every layer on the path).
template <typename TA, typename TW>
void launch_gemm_simt(const GemmParams& p, cudaStream_t st);

}  // namespace dsm
