// Split residual VQ encode as ONE kernel (quantization.rs:366-378 -> :219-229 -> :122-131,182-194).
//
// The per-level chain  scores = c2 - r.E^T -> first argmin -> r -= E[idx]  is strictly sequential over the 1 + 31
// codebooks; as separate launches (a tensor-core score GEMM + a select kernel per level) it costs 33 x ~20 us of
// mostly launch and pipeline latency at serving batch sizes.  Here a thread-block CLUSTER of 8 CTAs owns 16 streams for
// all levels and keeps their residuals on chip:
//   * CTA `rank` scores its 256 of the 2048 bins: E^T chunks [32 dims x 256 bins] stream through a 3-stage cp.async
//     ring (the ring never drains across levels: the next codebook does not depend on the argmin), a thread
//     accumulates 2 bins x 8 streams with f32 FMAs in dimension order;
//   * the per-stream minimum is a 64-bit key (order-preserving float bits << 32 | bin): warp shuffles, then one
//     shared-memory atomicMin per warp; the smallest bin wins ties (candle's argmin returns the first minimum);
//   * the 8 CTAs exchange their 16 keys through distributed shared memory (st.shared::cluster into every peer, one
//     barrier.cluster per level, key slots double-buffered by level parity), every CTA then knows the winner and
//     updates its own copy of the residual from the winning codebook row.
// CUDA cores, not tensor cores, on purpose: with 16 streams per cluster the MMA N dimension would be 16, where
// tcgen05 issue rate, not math, bounds the tile -- the f32 FMA pipe does the same work in the same time, exactly.
// MEASURED (B200, profiles/r02e): slower than the GEMM + select chain it was meant to replace -- codec step 3.36 vs
// 2.93 ms at 256 streams, equal at 64.  The 16 clusters each stream the whole codebook level (128 CTAs x 256 KB = 32 MB
// of L2 reads per level, the same 2 MB sixteen times) and a level costs ~34 us instead of the ~6 us the FMA count
// allows.  Kept opt-in (DSM_RVQ_FUSED_MAX) with its parity test; the chain stays the default.
#include "kernels.cuh"

namespace dsm {

namespace {

constexpr int RV_CS = 8;          // CTAs per cluster
constexpr int RV_SPC = 16;        // streams per cluster
constexpr int RV_BPC = 256;       // bins per CTA
constexpr int RV_KC = 32;         // dims per E^T chunk
constexpr int RV_STAGES = 3;
constexpr int RV_THREADS = 256;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_u64(uint32_t local_addr, uint32_t rank, unsigned long long v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(remote), "l"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// order-preserving map float -> uint32 (smaller float <=> smaller key); NaN sorts last
__device__ __forceinline__ uint32_t float_key(float f) {
  if (f != f) return 0xffffffffu;
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

}  // namespace

// embT [n_q][qdim][bins] (codebooks transposed at load), c2 [n_q][bins], emb [n_q][bins][qdim] (row gather),
// resid [B][2*qdim] = [first-group projection | rest-group projection] of the latent; codes [B][n_q].
__global__ void __launch_bounds__(RV_THREADS, 1)
rvq_fused_kernel(const float* __restrict__ embT, const float* __restrict__ c2, const float* __restrict__ emb,
                 const float* __restrict__ resid, int B, int n_q, int bins, int qdim, uint32_t* __restrict__ codes) {
  extern __shared__ __align__(16) uint8_t rv_smem[];
  float* e_ring = reinterpret_cast<float*>(rv_smem);                         // [STAGES][KC][BPC]
  float* r_first = e_ring + RV_STAGES * RV_KC * RV_BPC;                      // [qdim][SPC]  (level 0)
  float* r_rest = r_first + qdim * RV_SPC;                                   // [qdim][SPC]  (levels 1..)
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(r_rest + qdim * RV_SPC);   // [2][CS][SPC]
  unsigned long long* local_key = keys + 2 * RV_CS * RV_SPC;                 // [SPC]
  int* winner = reinterpret_cast<int*>(local_key + RV_SPC);                  // [SPC]

  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int s0 = (int)cluster_id_x() * RV_SPC;             // first stream of this cluster
  const int bin0 = (int)rank * RV_BPC;
  const int chunks_per_level = qdim / RV_KC;
  const int total_chunks = n_q * chunks_per_level;

  // E^T chunk `ci` (level ci / cpl, dims [KC * (ci % cpl), +KC)) of this CTA's bins -> ring slot ci % STAGES
  auto issue_chunk = [&](int ci) {
    if (ci < total_chunks) {
      const int q = ci / chunks_per_level, k0 = (ci % chunks_per_level) * RV_KC;
      const float* src = embT + ((long long)q * qdim + k0) * bins + bin0;
      float* dst = e_ring + (ci % RV_STAGES) * RV_KC * RV_BPC;
#pragma unroll
      for (int i = 0; i < RV_KC * RV_BPC / 4 / RV_THREADS; ++i) {
        const int idx = tid + RV_THREADS * i;            // 16-byte piece
        const int row = idx / (RV_BPC / 4), c4 = idx % (RV_BPC / 4);
        cp_async16(dst + row * RV_BPC + c4 * 4, src + (long long)row * bins + c4 * 4);
      }
    }
    cp_async_commit();                                   // (possibly empty) group: keeps the group count in step
  };
  issue_chunk(0);
  issue_chunk(1);

  // residuals, transposed to [dim][stream] so that a warp reads 4 streams of one dim as one broadcast float4
  for (int i = tid; i < RV_SPC * 2 * qdim; i += RV_THREADS) {
    const int s = i / (2 * qdim), c = i % (2 * qdim);
    const float v = (s0 + s < B) ? resid[(long long)(s0 + s) * 2 * qdim + c] : 0.f;
    if (c < qdim) r_first[c * RV_SPC + s] = v;
    else r_rest[(c - qdim) * RV_SPC + s] = v;
  }
  for (int i = tid; i < 2 * RV_CS * RV_SPC; i += RV_THREADS) keys[i] = ~0ULL;
  if (tid < RV_SPC) local_key[tid] = ~0ULL;
  __syncthreads();
  cluster_barrier();            // every CTA's key slots are initialised before anyone pushes into them

  // thread tile: bins {2*bp, 2*bp+1} x streams [8*sh, 8*sh+8)
  const int bp = tid & 127, sh = tid >> 7;
  int ci = 0;
  for (int q = 0; q < n_q; ++q) {
    float* r = q == 0 ? r_first : r_rest;
    float acc0[8], acc1[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) acc0[s] = 0.f, acc1[s] = 0.f;
    for (int c = 0; c < chunks_per_level; ++c, ++ci) {
      issue_chunk(ci + 2);
      cp_async_wait<2>();                                // chunk ci has landed (two younger groups may be in flight)
      __syncthreads();
      const float* e = e_ring + (ci % RV_STAGES) * RV_KC * RV_BPC + 2 * bp;
      const float* rk = r + (c * RV_KC) * RV_SPC + 8 * sh;
#pragma unroll 8
      for (int k = 0; k < RV_KC; ++k) {
        const float2 ev = *reinterpret_cast<const float2*>(e + k * RV_BPC);
        const float4 ra = *reinterpret_cast<const float4*>(rk + k * RV_SPC);
        const float4 rb = *reinterpret_cast<const float4*>(rk + k * RV_SPC + 4);
        acc0[0] = fmaf(ra.x, ev.x, acc0[0]), acc1[0] = fmaf(ra.x, ev.y, acc1[0]);
        acc0[1] = fmaf(ra.y, ev.x, acc0[1]), acc1[1] = fmaf(ra.y, ev.y, acc1[1]);
        acc0[2] = fmaf(ra.z, ev.x, acc0[2]), acc1[2] = fmaf(ra.z, ev.y, acc1[2]);
        acc0[3] = fmaf(ra.w, ev.x, acc0[3]), acc1[3] = fmaf(ra.w, ev.y, acc1[3]);
        acc0[4] = fmaf(rb.x, ev.x, acc0[4]), acc1[4] = fmaf(rb.x, ev.y, acc1[4]);
        acc0[5] = fmaf(rb.y, ev.x, acc0[5]), acc1[5] = fmaf(rb.y, ev.y, acc1[5]);
        acc0[6] = fmaf(rb.z, ev.x, acc0[6]), acc1[6] = fmaf(rb.z, ev.y, acc1[6]);
        acc0[7] = fmaf(rb.w, ev.x, acc0[7]), acc1[7] = fmaf(rb.w, ev.y, acc1[7]);
      }
      __syncthreads();                                   // the slot may be refilled (by the issue two chunks ahead)
    }
    // ---- c2 - dot, per-stream minimum over this CTA's bins (quantization.rs:128-129) ----
    const int j0 = bin0 + 2 * bp;
    const float2 cc = *reinterpret_cast<const float2*>(c2 + (long long)q * bins + j0);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const unsigned long long k0 = ((unsigned long long)float_key(cc.x - acc0[s]) << 32) | (uint32_t)j0;
      const unsigned long long k1 = ((unsigned long long)float_key(cc.y - acc1[s]) << 32) | (uint32_t)(j0 + 1);
      unsigned long long k = k0 < k1 ? k0 : k1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
        k = other < k ? other : k;
      }
      if (lane == 0) atomicMin(&local_key[8 * sh + s], k);
    }
    __syncthreads();
    // ---- exchange: my 16 keys into slot [parity][rank][stream] of every CTA of the cluster ----
    const int par = q & 1;
    if (tid < RV_CS * RV_SPC) {
      const int dst = tid / RV_SPC, s = tid % RV_SPC;
      st_cluster_u64(smem_addr(&keys[(par * RV_CS + rank) * RV_SPC + s]), (uint32_t)dst, local_key[s]);
    }
    cluster_barrier();
    if (tid < RV_SPC) {
      unsigned long long k = keys[(par * RV_CS + 0) * RV_SPC + tid];
#pragma unroll
      for (int rr = 1; rr < RV_CS; ++rr) {
        const unsigned long long o = keys[(par * RV_CS + rr) * RV_SPC + tid];
        k = o < k ? o : k;
      }
      int idx = (int)(uint32_t)(k & 0xffffffffu);
      if ((uint32_t)(k >> 32) == 0xffffffffu || idx < 0 || idx >= bins) idx = 0;   // all-NaN row: candle's argmin returns 0
      winner[tid] = idx;
      local_key[tid] = ~0ULL;
      if (rank == 0 && s0 + tid < B) codes[(long long)(s0 + tid) * n_q + q] = (uint32_t)idx;
    }
    __syncthreads();
    // ---- residual update r <- r - E[idx] (quantization.rs:226); level 0 leaves the rest-group residual untouched ----
    if (q > 0 && q + 1 < n_q) {
      // a thread takes one dim of all 16 winning rows: 16 independent loads in flight (one L2 round trip per level on
      // the critical path instead of 16), then one 64-byte row of the transposed residual
      const float* eq = emb + (long long)q * bins * qdim;
      for (int d = tid; d < qdim; d += RV_THREADS) {
        float ev[RV_SPC];
#pragma unroll
        for (int s = 0; s < RV_SPC; ++s) ev[s] = __ldg(eq + (long long)winner[s] * qdim + d);
        float4* row = reinterpret_cast<float4*>(r_rest + d * RV_SPC);
#pragma unroll
        for (int s4 = 0; s4 < RV_SPC / 4; ++s4) {
          float4 v = row[s4];
          v.x -= ev[4 * s4], v.y -= ev[4 * s4 + 1], v.z -= ev[4 * s4 + 2], v.w -= ev[4 * s4 + 3];
          row[s4] = v;
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  cluster_barrier();            // no CTA exits while a peer may still push into its shared memory
}

size_t rvq_fused_smem(int qdim) {
  return (size_t)RV_STAGES * RV_KC * RV_BPC * 4 + (size_t)2 * qdim * RV_SPC * 4 + (size_t)(2 * RV_CS * RV_SPC + RV_SPC) * 8 + RV_SPC * 4;
}

bool rvq_fused_supported(int bins, int qdim) { return bins == RV_CS * RV_BPC && qdim % RV_KC == 0 && qdim >= RV_KC; }

void launch_rvq_fused(const float* embT, const float* c2, const float* emb, const float* resid, int B, int n_q, int bins,
                      int qdim, uint32_t* codes, cudaStream_t st) {
  DSM_CHECK(rvq_fused_supported(bins, qdim), DSM_ERR_UNSUPPORTED, "rvq_fused: bins=%d qdim=%d", bins, qdim);
  const size_t smem = rvq_fused_smem(qdim);
  static PerDeviceOnce attr;
  attr.run([&] { DSM_CUDA(cudaFuncSetAttribute(rvq_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); });
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(ceil_div(B, RV_SPC) * RV_CS)), cfg.blockDim = dim3(RV_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  unsigned na = 0;
  attrs[na].id = cudaLaunchAttributeClusterDimension;
  attrs[na].val.clusterDim.x = RV_CS, attrs[na].val.clusterDim.y = 1, attrs[na].val.clusterDim.z = 1;
  ++na;
  if (pdl_enabled()) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attrs, cfg.numAttrs = na;
  DSM_CUDA(cudaLaunchKernelEx(&cfg, rvq_fused_kernel, embT, c2, emb, resid, B, n_q, bins, qdim, codes));
  check_launch("rvq_fused");
}

}  // namespace dsm
