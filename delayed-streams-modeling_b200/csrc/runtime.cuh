// Host-side runtime pieces shared by the Mimi / LM / ASR handles: device arena, weight table access,
// host<->device staging, CUDA-graph capture of a launch plan.
#pragma once

#include <functional>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace dsm {

// Owns every cudaMalloc of a handle; freed with the handle.
class Arena {
 public:
  Arena() = default;
  Arena(const Arena&) = delete;
  Arena& operator=(const Arena&) = delete;
  ~Arena();
  template <typename T>
  T* alloc(size_t n, bool zero = true) {
    return static_cast<T*>(alloc_bytes(n * sizeof(T), zero));
  }
  template <typename T>
  T* upload(const std::vector<T>& host) {
    T* d = alloc<T>(host.size(), false);
    DSM_CUDA(cudaMemcpy(d, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
  }
  void* alloc_bytes(size_t bytes, bool zero);
  void* alloc_pinned(size_t bytes);
  size_t total_bytes() const { return total_; }

 private:
  std::vector<void*> dev_, pinned_;
  size_t total_ = 0;
};

// Name -> tensor view over the caller's dsm_weights table.
class WeightTable {
 public:
  explicit WeightTable(const dsm_weights* w);
  bool has(const std::string& name) const { return map_.count(name) != 0; }
  // Fetches a tensor as float32 (widening bf16 / f16), checking the element count.
  std::vector<float> f32(const std::string& name, size_t expect_numel) const;
  const dsm_tensor* find(const std::string& name) const;

 private:
  std::map<std::string, const dsm_tensor*> map_;
};

bool is_device_pointer(const void* p);

// f32 host vector -> device bf16 (round-to-nearest-even on the device).
__nv_bfloat16* upload_bf16(Arena& arena, const std::vector<float>& host);
// Same, from a raw table entry without a host-side f32 copy (used for the big LM matrices).
__nv_bfloat16* upload_tensor_bf16(Arena& arena, const dsm_tensor* t, size_t expect_numel, cudaStream_t st);
// Seeded device-side init: uniform with the given standard deviation (benchmarks without checkpoints).
void fill_random_f32(float* dst, size_t n, float std, float mean, uint64_t seed, cudaStream_t st);
void fill_random_bf16(__nv_bfloat16* dst, size_t n, float std, uint64_t seed, cudaStream_t st);
void fill_const_f32(float* dst, size_t n, float v, cudaStream_t st);
void fill_u8(uint8_t* dst, size_t n, uint8_t v, cudaStream_t st);
void fill_u32(uint32_t* dst, size_t n, uint32_t v, cudaStream_t st);

// A launch plan: closures that enqueue kernels on a stream.  Replayed eagerly or through a CUDA graph.
using Plan = std::vector<std::function<void(cudaStream_t)>>;
inline void run_plan(const Plan& plan, cudaStream_t st) {
  for (auto& f : plan) f(st);
}

class GraphExec {
 public:
  ~GraphExec();
  bool ready() const { return exec_ != nullptr; }
  // Captures `body` (which may fork onto other streams and must join back) on `st`.
  void capture(cudaStream_t st, const std::function<void()>& body);
  void launch(cudaStream_t st);
  void reset();
  size_t num_nodes() const { return nodes_; }

 private:
  cudaGraph_t graph_ = nullptr;
  cudaGraphExec_t exec_ = nullptr;
  size_t nodes_ = 0;
  long long launches_ = 0;  // kernel launches inside one replay
};

// SM partitioning of the frame graph's two branches (CUDA green contexts).  The LM branch is HBM-bound (ring
// attention at the copy peak) and the codec branch SM-bound (tensor / FMA work); sharing the SMs they queue for whole
// SMs and the step costs nearly the sum of both.  Confined to disjoint SM sets they overlap: 256 streams 7.45 -> 6.5 ms,
// 512 streams 14.05 -> 12.1 ms, stt-2.6b x 128 7.54 -> 7.09 ms with 96 + 52 SMs (measured); at
// 2,000 streams, where the LM's projections are tensor-bound themselves, it loses (37.2 -> 39.0+ ms).
//   sm_split_choice(B, kv_8bit): SMs for the LM branch of an ASR state with B streams, 0 = no partition.
//   DSM_SM_SPLIT=<n> forces n (0 = never); unset, the measured rule applies.
int sm_split_choice(int batch, bool kv_8bit);
struct SmPartitionInfo {
  int lm_sms = 0, codec_sms = 0;
};
// Creates (once per device and size) the two green contexts and returns a non-blocking stream on each.
SmPartitionInfo create_partition_streams(int device, int lm_sms, cudaStream_t* lm_stream, cudaStream_t* codec_stream);
bool graphs_enabled();  // false when DSM_NO_GRAPH=1 or dsm_set_graphs(0)
void set_graphs_enabled(int v);
const std::string& last_error();

}  // namespace dsm
