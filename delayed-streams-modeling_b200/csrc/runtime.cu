#include "runtime.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

namespace dsm {

thread_local long long g_launch_count = 0;
static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }
const std::string& last_error() { return g_last_error; }

// ---- Arena -------------------------------------------------------------------------------------
Arena::~Arena() {
  for (void* p : dev_) cudaFree(p);
  for (void* p : pinned_) cudaFreeHost(p);
}

void* Arena::alloc_bytes(size_t bytes, bool zero) {
  if (bytes == 0) bytes = 16;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    fail(DSM_ERR_OOM, "cudaMalloc of %zu bytes failed: %s (handle already holds %zu bytes)", bytes,
         cudaGetErrorString(e), total_);
  }
  dev_.push_back(p);
  total_ += bytes;
  if (zero) DSM_CUDA(cudaMemset(p, 0, bytes));
  return p;
}

void* Arena::alloc_pinned(size_t bytes) {
  void* p = nullptr;
  DSM_CUDA(cudaMallocHost(&p, bytes ? bytes : 16));
  pinned_.push_back(p);
  return p;
}

// ---- WeightTable -------------------------------------------------------------------------------
WeightTable::WeightTable(const dsm_weights* w) {
  if (!w) return;
  for (int64_t i = 0; i < w->count; ++i) {
    const dsm_tensor& t = w->tensors[i];
    if (t.name && t.data) map_[t.name] = &t;
  }
}

static size_t numel(const dsm_tensor* t) {
  size_t n = 1;
  for (int i = 0; i < t->ndim; ++i) n *= (size_t)t->shape[i];
  return n;
}

const dsm_tensor* WeightTable::find(const std::string& name) const {
  auto it = map_.find(name);
  if (it == map_.end()) fail(DSM_ERR_MISSING_WEIGHT, "missing weight tensor '%s'", name.c_str());
  return it->second;
}

static inline float half_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ff, u;
  if (exp == 0) {
    if (man == 0) u = sign;
    else {
      int e = -1;
      do { man <<= 1; ++e; } while (!(man & 0x400));
      u = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ff) << 13);
    }
  } else if (exp == 31) u = sign | 0x7f800000u | (man << 13);
  else u = sign | ((exp + 112) << 23) | (man << 13);
  float f;
  memcpy(&f, &u, 4);
  return f;
}

std::vector<float> WeightTable::f32(const std::string& name, size_t expect_numel) const {
  const dsm_tensor* t = find(name);
  const size_t n = numel(t);
  DSM_CHECK(n == expect_numel, DSM_ERR_SHAPE, "weight '%s' has %zu elements, expected %zu", name.c_str(), n, expect_numel);
  std::vector<float> out(n);
  if (t->dtype == DSM_F32) memcpy(out.data(), t->data, n * 4);
  else if (t->dtype == DSM_BF16) {
    const uint16_t* s = static_cast<const uint16_t*>(t->data);
    for (size_t i = 0; i < n; ++i) {
      uint32_t u = (uint32_t)s[i] << 16;
      memcpy(&out[i], &u, 4);
    }
  } else if (t->dtype == DSM_F16) {
    const uint16_t* s = static_cast<const uint16_t*>(t->data);
    for (size_t i = 0; i < n; ++i) out[i] = half_to_float(s[i]);
  } else fail(DSM_ERR_INVALID_ARG, "weight '%s': unsupported dtype %d", name.c_str(), t->dtype);
  return out;
}

bool is_device_pointer(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---- conversions / fills -------------------------------------------------------------------------
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i]);
}
__global__ void f16_to_bf16_kernel(const __half* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = __float2bfloat16_rn(__half2float(src[i]));
}

__nv_bfloat16* upload_bf16(Arena& arena, const std::vector<float>& host) {
  dsm_tensor t{};
  t.name = "<host>";
  t.dtype = DSM_F32;
  t.ndim = 1;
  t.shape[0] = (int64_t)host.size();
  t.data = host.data();
  return upload_tensor_bf16(arena, &t, host.size(), 0);
}

__nv_bfloat16* upload_tensor_bf16(Arena& arena, const dsm_tensor* t, size_t expect_numel, cudaStream_t st) {
  const size_t n = numel(t);
  DSM_CHECK(n == expect_numel, DSM_ERR_SHAPE, "weight '%s' has %zu elements, expected %zu", t->name, n, expect_numel);
  __nv_bfloat16* dst = arena.alloc<__nv_bfloat16>(n, false);
  if (t->dtype == DSM_BF16) {
    DSM_CUDA(cudaMemcpy(dst, t->data, n * 2, cudaMemcpyHostToDevice));
    return dst;
  }
  // stage through a bounded scratch buffer and round on the device
  const size_t chunk = std::min<size_t>(n, (size_t)32 << 20);
  const size_t esz = t->dtype == DSM_F32 ? 4 : 2;
  void* scratch = nullptr;
  DSM_CUDA(cudaMalloc(&scratch, chunk * esz));
  try {
    for (size_t o = 0; o < n; o += chunk) {
      const size_t c = std::min(chunk, n - o);
      DSM_CUDA(cudaMemcpyAsync(scratch, static_cast<const char*>(t->data) + o * esz, c * esz, cudaMemcpyHostToDevice, st));
      if (t->dtype == DSM_F32)
        f32_to_bf16_kernel<<<1184, 256, 0, st>>>(static_cast<const float*>(scratch), dst + o, c);
      else if (t->dtype == DSM_F16)
        f16_to_bf16_kernel<<<1184, 256, 0, st>>>(static_cast<const __half*>(scratch), dst + o, c);
      else fail(DSM_ERR_INVALID_ARG, "weight '%s': unsupported dtype %d", t->name, t->dtype);
      DSM_CUDA(cudaGetLastError());
      DSM_CUDA(cudaStreamSynchronize(st));
    }
  } catch (...) {
    cudaFree(scratch);
    throw;
  }
  cudaFree(scratch);
  return dst;
}

__device__ __forceinline__ uint32_t hash_u64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return (uint32_t)x;
}
// uniform in [-sqrt(3), sqrt(3)): unit variance
__device__ __forceinline__ float unit_uniform(uint64_t seed, size_t i) {
  const uint32_t h = hash_u64(seed * 0x9e3779b97f4a7c15ULL + i);
  return ((float)(h >> 8) * (1.f / 8388608.f) - 1.f) * 1.7320508f;
}
__global__ void fill_random_f32_kernel(float* dst, size_t n, float std, float mean, uint64_t seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = mean + std * unit_uniform(seed, i);
}
__global__ void fill_random_bf16_kernel(__nv_bfloat16* dst, size_t n, float std, uint64_t seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = __float2bfloat16_rn(std * unit_uniform(seed, i));
}
template <typename T>
__global__ void fill_const_kernel(T* dst, size_t n, T v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = v;
}
static unsigned fill_grid(size_t n) { return (unsigned)std::min<size_t>(ceil_div<size_t>(n, 256), 148 * 8); }

void fill_random_f32(float* dst, size_t n, float std, float mean, uint64_t seed, cudaStream_t st) {
  if (!n) return;
  fill_random_f32_kernel<<<fill_grid(n), 256, 0, st>>>(dst, n, std, mean, seed);
  DSM_CUDA(cudaGetLastError());
}
void fill_random_bf16(__nv_bfloat16* dst, size_t n, float std, uint64_t seed, cudaStream_t st) {
  if (!n) return;
  fill_random_bf16_kernel<<<fill_grid(n), 256, 0, st>>>(dst, n, std, seed);
  DSM_CUDA(cudaGetLastError());
}
void fill_const_f32(float* dst, size_t n, float v, cudaStream_t st) {
  if (!n) return;
  fill_const_kernel<float><<<fill_grid(n), 256, 0, st>>>(dst, n, v);
  DSM_CUDA(cudaGetLastError());
}
void fill_u8(uint8_t* dst, size_t n, uint8_t v, cudaStream_t st) {
  if (!n) return;
  fill_const_kernel<uint8_t><<<fill_grid(n), 256, 0, st>>>(dst, n, v);
  DSM_CUDA(cudaGetLastError());
}
void fill_u32(uint32_t* dst, size_t n, uint32_t v, cudaStream_t st) {
  if (!n) return;
  fill_const_kernel<uint32_t><<<fill_grid(n), 256, 0, st>>>(dst, n, v);
  DSM_CUDA(cudaGetLastError());
}

// ---- CUDA graph -----------------------------------------------------------------------------------
static int g_graphs_override = -1;   // -1: follow DSM_NO_GRAPH, 0/1: forced by dsm_set_graphs
void set_graphs_enabled(int v) { g_graphs_override = v; }
bool pdl_enabled() {
  // Measured on B200 at 256 streams (profiles/r02b_*): no gain for either branch of the frame graph alone (LM 5.27 vs
  // 5.25 ms, codec 2.92 vs 2.84 ms) and a LOSS when they run side by side (7.68 vs 7.18 ms): early-resident CTAs
  // spinning in griddepcontrol.wait hold the SM's registers / shared memory that the other branch's kernels could use.
  // So it is opt-in (DSM_PDL=1).
  static const bool on = [] { const char* e = getenv("DSM_PDL"); return e && e[0] == '1'; }();
  return on;
}

// ---- SM partitions ------------------------------------------------------------------------------
namespace {
struct SmPartition {
  bool ok = false;
  CUgreenCtx ctx[2] = {nullptr, nullptr};
  int sms[2] = {0, 0};
};
template <typename F>
F driver_fn(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
    fail(DSM_ERR_UNSUPPORTED, "driver entry point %s is not available (SM partitions need CUDA green contexts)", name);
  return reinterpret_cast<F>(p);
}
void cu_check(CUresult r, const char* what) {
  if (r != CUDA_SUCCESS) fail(DSM_ERR_CUDA, "%s failed with CUresult %d", what, (int)r);
}
SmPartition& sm_partition(int device, int lm_sms) {
  static std::mutex mu;
  static std::map<std::pair<int, int>, SmPartition> parts;
  std::lock_guard<std::mutex> l(mu);
  SmPartition& P = parts[{device, lm_sms}];
  if (P.ok) return P;
  DSM_CUDA(cudaSetDevice(device));
  DSM_CUDA(cudaFree(nullptr));                 // the primary context must exist
  auto dev_get = driver_fn<CUresult (*)(CUdevice*, int)>("cuDeviceGet");
  auto get_res = driver_fn<CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType)>("cuDeviceGetDevResource");
  auto split = driver_fn<CUresult (*)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int)>(
      "cuDevSmResourceSplitByCount");
  auto gen = driver_fn<CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned int)>("cuDevResourceGenerateDesc");
  auto create = driver_fn<CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int)>("cuGreenCtxCreate");
  CUdevice dev;
  cu_check(dev_get(&dev, device), "cuDeviceGet");
  CUdevResource all, part[1], rest;
  cu_check(get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM), "cuDeviceGetDevResource");
  unsigned int groups = 1;
  cu_check(split(part, &groups, &all, &rest, 0, (unsigned)lm_sms), "cuDevSmResourceSplitByCount");
  DSM_CHECK(groups == 1 && rest.sm.smCount > 0, DSM_ERR_INVALID_ARG, "an LM partition of %d SMs leaves none for the codec branch (%u in all)",
            lm_sms, all.sm.smCount);
  CUdevResource* src[2] = {&part[0], &rest};
  for (int i = 0; i < 2; ++i) {
    CUdevResourceDesc desc;
    cu_check(gen(&desc, src[i], 1), "cuDevResourceGenerateDesc");
    cu_check(create(&P.ctx[i], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM), "cuGreenCtxCreate");
    P.sms[i] = (int)src[i]->sm.smCount;
  }
  P.ok = true;
  return P;
}
}  // namespace

int sm_split_choice(int batch, bool kv_8bit) {
  static const int forced = [] { const char* e = getenv("DSM_SM_SPLIT"); return e ? atoi(e) : -1; }();
  if (forced >= 0) return forced;
  // Measured (tools/gpu_run_ab.sh, _ac, _ad, _ao): with the bf16 ring 96 + 52 SMs wins 5-14 % at every
  // batch tried from 1 to 1,536 streams (the memory ceiling is ~1,700); with the int8 ring (half the attention time) it
  // wins 2-4 % at 256 and 512, nothing at 1,024, and loses 5-8 % at 2,000, where the LM's projections are tensor-bound
  // and want every SM.  Outside the measured range the branches share the device.
  // Nsight Compute / compute-sanitizer cannot instrument kernels launched into green contexts ("Failed to prepare
  // kernel for profiling"): under an injected tool the branches share the device, so per-kernel captures keep working.
  static const bool tool = [] {
    const bool t = getenv("CUDA_INJECTION64_PATH") || getenv("NV_NSIGHT_INJECTION_PORT_BASE") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR");
    if (t) fprintf(stderr, "dsm: profiling / sanitizer tool attached: SM partition off (set DSM_SM_SPLIT=<n> to force it)\n");
    return t;
  }();
  if (tool) return 0;
  return (!kv_8bit || batch <= 1024) ? 96 : 0;
}

SmPartitionInfo create_partition_streams(int device, int lm_sms, cudaStream_t* lm_stream, cudaStream_t* codec_stream) {
  SmPartition& P = sm_partition(device, lm_sms);
  auto mk = driver_fn<CUresult (*)(CUstream*, CUgreenCtx, unsigned int, int)>("cuGreenCtxStreamCreate");
  cudaStream_t* out[2] = {lm_stream, codec_stream};
  for (int i = 0; i < 2; ++i) {
    CUstream cs = nullptr;
    cu_check(mk(&cs, P.ctx[i], CU_STREAM_NON_BLOCKING, 0), "cuGreenCtxStreamCreate");
    *out[i] = reinterpret_cast<cudaStream_t>(cs);
  }
  SmPartitionInfo info;
  info.lm_sms = P.sms[0], info.codec_sms = P.sms[1];
  return info;
}

bool graphs_enabled() {
  if (g_graphs_override >= 0) return g_graphs_override != 0;
  const char* e = getenv("DSM_NO_GRAPH");
  return !(e && e[0] == '1');
}

GraphExec::~GraphExec() { reset(); }

void GraphExec::reset() {
  if (exec_) cudaGraphExecDestroy(exec_);
  if (graph_) cudaGraphDestroy(graph_);
  exec_ = nullptr;
  graph_ = nullptr;
  nodes_ = 0;
}

void GraphExec::capture(cudaStream_t st, const std::function<void()>& body) {
  reset();
  DSM_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const long long before = g_launch_count;
  try {
    body();
  } catch (...) {
    cudaGraph_t g = nullptr;
    cudaStreamEndCapture(st, &g);
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    throw;
  }
  launches_ = g_launch_count - before;
  g_launch_count = before;  // captured launches are counted when the graph is replayed
  DSM_CUDA(cudaStreamEndCapture(st, &graph_));
  DSM_CUDA(cudaGraphGetNodes(graph_, nullptr, &nodes_));
  DSM_CUDA(cudaGraphInstantiate(&exec_, graph_, 0));
}

void GraphExec::launch(cudaStream_t st) {
  DSM_CHECK(exec_ != nullptr, DSM_ERR_INVALID_ARG, "graph launched before capture");
  DSM_CUDA(cudaGraphLaunch(exec_, st));
  g_launch_count += launches_;
}

}  // namespace dsm
