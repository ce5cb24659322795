// tcgen05 / TMEM / TMA GEMM for the LM's dense projections (sm_100a):  C[m, n] = sum_k X[m, k] * W[n, k]
// with bf16 operands and f32 accumulation.
//
// Batched decode makes these GEMMs weight-streaming (M = streams is small, W is read once per step),
// so the tile is "swap-AB": the WEIGHT rows fill the 128-row MMA M dimension (TMEM lanes) and the
// streams are the MMA N dimension (TMEM columns).  One CTA = 128 weight rows x BMX streams x a K slice:
//   warp 0 / one lane : TMA producer (cp.async.bulk.tensor, 128B-swizzled K-major tiles, mbarrier tx)
//   warp 1 / one lane : tcgen05.mma issuer (kind::f16, A and B from shared memory, D in TMEM)
//   warp 2            : TMEM alloc / dealloc
//   warps 0-3         : epilogue -- tcgen05.ld of the warp's 32 TMEM lanes, fused op, coalesced stores
// Epilogues: plain f32 store (QKV, text head), residual accumulate x += acc (out-proj, MLP-out; with
// split-K every split stores its partial tile and the kernel that consumes the result adds the partials in split
// order, so results are bit-reproducible from run to run), and the gated-MLP pair
// silu(gate) * value -> bf16, where the CTA accumulates gate rows [j0, j0+128) and value rows
// [hidden + j0, ...) side by side in TMEM (batched_transformer.rs:170-176).
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace dsm {

enum TcEpilogue { TC_STORE_F32 = 0, TC_ACCUM_F32 = 1, TC_GATE_BF16 = 2 };

struct TcGemm {
  CUtensorMap map_w;   // [N, K] bf16, box 64 x 128, 128B swizzle
  CUtensorMap map_x;   // [M, K] bf16, box 64 x BMX
  int M = 0, N = 0, K = 0;
  int bmx = 0;         // streams per tile: 64, 128 or 256
  int nacc = 1;        // weight sub-tiles (TMEM accumulators) per CTA
  int split_k = 1;
  int mode = TC_STORE_F32;
  int hidden = 0;      // TC_GATE_BF16: N = 2 * hidden
  float* out_f32 = nullptr;            // STORE: [M, N];  ACCUM: x [M, N] (+=)   (split_k == 1, or after tc_reduce_partials)
  __nv_bfloat16* out_bf16 = nullptr;   // GATE: [M, hidden]
  // split_k > 1: the kernel writes partial sums to partials [split_k][M][N] (partial_bytes, set by prepare; the
  // caller binds the memory) and the CONSUMER of the result adds them in split order -- see lm.cu
  size_t partial_bytes = 0;
  float* partials = nullptr;
};

// Builds the tensor maps and picks the tile / split-K so the grid covers the 148 SMs.
// sms_budget: the SMs the GEMM may count on (tile shape and split-K are chosen to fill them)
void tc_gemm_prepare(TcGemm& g, const __nv_bfloat16* W, const __nv_bfloat16* X, int M, int N, int K, int mode,
                     float* out_f32, __nv_bfloat16* out_bf16, int sms_budget = 148);
void tc_gemm_launch(const TcGemm& g, cudaStream_t st);
void tc_reduce_partials(const TcGemm& g, cudaStream_t st);   // out (+)= sum of the partial planes, in split order
bool tc_gemm_supported(int M, int N, int K, int mode);

}  // namespace dsm
