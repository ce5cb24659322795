// Shared helpers for the dsm_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>

#include "../../include/dsm_b200.h"

namespace dsm {

// ---- error plumbing: C++ exceptions inside, status codes at the ABI -------------------------
struct Error : std::runtime_error {
  dsm_status code;
  Error(dsm_status c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg);

[[noreturn]] inline void fail(dsm_status code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error(code, buf);
}

#define DSM_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      ::dsm::fail(_e == cudaErrorMemoryAllocation ? DSM_ERR_OOM : DSM_ERR_CUDA,                \
                  "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define DSM_CHECK(cond, code, ...) \
  do {                             \
    if (!(cond)) ::dsm::fail(code, __VA_ARGS__); \
  } while (0)

// Launch-count bookkeeping (bench.py reports "gpu_launches").
extern thread_local long long g_launch_count;
#define DSM_LAUNCHED() (++::dsm::g_launch_count)

inline void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fail(DSM_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  DSM_LAUNCHED();
}

// ---- programmatic dependent launch ---------------------------------------------------------------
// With DSM_PDL=1 every step kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization; each begins with
//     pdl_launch_dependents();   // the next kernel of the stream / graph branch may be scheduled once all of this
//                                // grid's CTAs have started
//     pdl_wait();                // ... but nothing of THIS grid touches global memory before its predecessors have
//                                // completed and flushed
// so a kernel's launch latency, CTA rasterisation and (for the tensor-core GEMMs) barrier / TMEM / descriptor setup
// overlap the tail of the kernel before it, inside the captured frame graph too.  A kernel launched through
// launch_k() MUST call pdl_wait() before its first global access.  Off by default: measured slower on the two-branch
// frame graph (see pdl_enabled() in runtime.cu); without the launch attribute both instructions are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (pdl_enabled()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
  }
  DSM_CUDA(cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...));
}

// Row addressing shared by every row-wise kernel: logical row m = b * T + t lives at
// base + b * stride_b + t * stride_t (elements).  Lets a layer read a window of, or write into, a
// neighbour's state-carrying buffer without copies.
struct RowMap {
  int T;
  long long stride_b;
  long long stride_t;
  __host__ __device__ long long off(int m) const {
    int b = m / T, t = m - b * T;
    return b * stride_b + t * stride_t;
  }
};
inline RowMap dense_rows(int T, long long row_len) { return RowMap{T, (long long)T * row_len, row_len}; }

// cudaFuncSetAttribute is per DEVICE: remembers on which devices a kernel has been configured (a handle may live on
// any GPU of the process).  The configuring call is idempotent, so a race between two threads is benign.
struct PerDeviceOnce {
  unsigned long long done = 0;
  template <typename F>
  void run(F&& f) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ULL << (dev & 63);
    if (__atomic_load_n(&done, __ATOMIC_ACQUIRE) & bit) return;
    f();
    __atomic_fetch_or(&done, bit, __ATOMIC_RELEASE);
  }
};

template <typename T>
inline T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}

// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// ELU(alpha = 1) as the reference computes it: exp(x) - 1 for x <= 0 (candle Activation::Elu), not expm1
__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expf(x) - 1.f; }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.f + expf(-x)); }

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// Exact 3-way split of an f32 by truncation: x = h + m + l with bf16 h, m, l (8 + 8 + 8 significant bits).  The
// tcgen05 codec GEMM takes activations in this form (gemm_tc6.cu); producers that feed its skinny GEMMs store it.
__device__ __forceinline__ void split3_bf16(float x, uint16_t& h, uint16_t& m, uint16_t& l) {
  const uint32_t u = __float_as_uint(x);
  h = (uint16_t)(u >> 16);
  const float r1 = x - __uint_as_float(u & 0xffff0000u);
  const uint32_t v = __float_as_uint(r1);
  m = (uint16_t)(v >> 16);
  const float r2 = r1 - __uint_as_float(v & 0xffff0000u);
  l = (uint16_t)(__float_as_uint(r2) >> 16);
}
__device__ __forceinline__ void store_planes4(__nv_bfloat16* base, long long plane_stride, long long off, const float4& x) {
  uint16_t h[4], m[4], l[4];
  split3_bf16(x.x, h[0], m[0], l[0]);
  split3_bf16(x.y, h[1], m[1], l[1]);
  split3_bf16(x.z, h[2], m[2], l[2]);
  split3_bf16(x.w, h[3], m[3], l[3]);
  *reinterpret_cast<uint2*>(base + off) = make_uint2(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16));
  *reinterpret_cast<uint2*>(base + plane_stride + off) = make_uint2(m[0] | ((uint32_t)m[1] << 16), m[2] | ((uint32_t)m[3] << 16));
  *reinterpret_cast<uint2*>(base + 2 * plane_stride + off) = make_uint2(l[0] | ((uint32_t)l[1] << 16), l[2] | ((uint32_t)l[3] << 16));
}

}  // namespace dsm
