// f32-accurate tensor-core GEMM for the codec (3-way bf16 split, 6 tcgen05 MMAs); see gemm_tc6.cu.
#pragma once

#include <cuda.h>

#include <vector>

#include "gemm_simt.cuh"
#include "runtime.cuh"

namespace dsm {

// Scratch shared by every split-K GEMM of one stream (they run one after another): partial-sum tiles and
// per-tile arrival counters (zero between launches: the finishing CTA resets its counter).
struct Tc6Workspace {
  float* ws = nullptr;
  size_t ws_bytes = 0;
  int* counters = nullptr;
  int n_counters = 0;
};
Tc6Workspace tc6_make_workspace(Arena& arena);

struct Tc6Gemm {
  CUtensorMap map_a;   // f32 [B][T_out][K] window view of the activation buffer
  CUtensorMap map_w;   // bf16 [3][N][K] pre-split weight planes
  int K = 0, N = 0, T_out = 0, B = 0;
  int t_sub = 0, b_sub = 0, bn = 0, nj = 0, pm = 1, split_k = 1, acc_bufs = 2;
  int sms = 148;       // persistent grid size cap: the SMs of the branch's partition (runtime.cuh)
  bool presplit = false;   // A arrives as bf16 planes (GemmParams::a_planes): TMA feeds the operand buffers directly
  float* ws = nullptr;
  int* counters = nullptr;
  GemmParams ep;       // the problem as the CUDA-core GEMM would take it (epilogue fields are reused)
};

// True when the problem fits the kernel's constraints (N, K multiples of 32; 16-byte aligned strides).
bool tc6_supported(const GemmParams& g);
// Splits an f32 [N, K] weight matrix into three bf16 planes on the device ([3][N][K]).
__nv_bfloat16* tc6_split_weights(Arena& arena, const std::vector<float>& w, int N, int K);
// sms: the SMs the kernel may count on (all 148, or the codec partition of an ASR state: runtime.cuh) -- it sizes the
// persistent grid, the column grouping and the split-K.  May be called again on a prepared GEMM to re-tune it.
void tc6_prepare(Tc6Gemm& t, const GemmParams& g, const __nv_bfloat16* w_planes, const Tc6Workspace& wsp = Tc6Workspace(),
                 int sms = 148);
void tc6_launch(const Tc6Gemm& t, cudaStream_t st);

}  // namespace dsm
