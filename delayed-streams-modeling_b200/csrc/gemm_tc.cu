#include "gemm_tc.cuh"

#include <mutex>

namespace dsm {

// ================================================================================================
// PTX wrappers (sm_100a)
// ================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp; the compiler then keeps tcgen05 / TMA operands in uniform registers instead of
// wrapping every instruction in a lane-election loop (what `lane == 0` produces).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> f32, M = 128
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128B-swizzled operand tile whose rows are exactly 128 bytes (64 bf16): 8-row groups are
// 1024 bytes apart (SBO), LBO is unused for swizzled K-major layouts (encoded as 1), descriptor
// version 1 (Blackwell), layout type 2 = SWIZZLE_128B.  (cute/arch/mma_sm100_desc.hpp SmemDescriptor.)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;              // leading byte offset (16 B units)
  d |= (uint64_t)(1024 >> 4) << 32;    // stride byte offset
  d |= (uint64_t)1 << 46;              // version
  d |= (uint64_t)2 << 61;              // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ================================================================================================
// kernel
// ================================================================================================
constexpr int TC_BN = 128;      // weight rows per CTA (UMMA M)
constexpr int TC_BK = 64;       // bf16 per k-block = one 128-byte swizzle row
constexpr int TC_UK = 16;       // UMMA K for 16-bit inputs
constexpr int TC_MAX_SPLIT = 16;

struct TcParams {
  int M, N, K;
  int kb_total;      // K / TC_BK
  int kb_per_split;  // k-blocks per blockIdx.z (the last split may get fewer)
  int hidden;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  float* partials;   // split-K: [gridDim.z][M][N]
};

// NACC weight sub-tiles (accumulators) per CTA: the gate/value pair, or two consecutive 128-row tiles
// ("wide" CTAs halve how often the activation slab is pulled through L2 per weight byte).
template <int BMX, int MODE, int NACC, int STAGES>
__global__ void __launch_bounds__(128, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const TcParams p) {
  static_assert(MODE != TC_GATE_BF16 || NACC == 2, "the gated epilogue needs the gate and value accumulators");
  constexpr int A_BYTES = TC_BN * TC_BK * 2;          // 16 KB
  constexpr int B_BYTES = BMX * TC_BK * 2;
  constexpr int STAGE_BYTES = NACC * A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = NACC * BMX < 32 ? 32 : NACC * BMX;
  static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS <= 512, "TMEM columns must be a power of two <= 512");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;
  // weight-row tile: GATE -> gate rows [n0, n0+128) and value rows hidden + the same; otherwise NACC
  // consecutive 128-row tiles starting at n0
  const int n0 = blockIdx.x * TC_BN * (MODE == TC_GATE_BF16 ? 1 : NACC);
  const int second = MODE == TC_GATE_BF16 ? p.hidden + n0 : n0 + TC_BN;
  const int m0 = blockIdx.y * BMX;         // stream tile
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int num_kb = min(p.kb_per_split, p.kb_total - kb_begin);
  const int k0 = kb_begin * TC_BK;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_x);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above is on-chip setup; the operands and the output belong to the kernels before this one
  pdl_wait();

  if (warp == 0 && elect_one()) {
    // ---------------- TMA producer ----------------
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      mbar_expect_tx(&full_bar[s], STAGE_BYTES);
      uint8_t* st = smem + s * STAGE_BYTES;
      const int kc = k0 + kb * TC_BK;
      tma_load_2d(st, &map_w, kc, n0, &full_bar[s]);
      if (NACC == 2) tma_load_2d(st + A_BYTES, &map_w, kc, second, &full_bar[s]);
      tma_load_2d(st + NACC * A_BYTES, &map_x, kc, m0, &full_bar[s]);
    }
  } else if (warp == 1 && elect_one()) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = make_idesc(BMX);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
      const uint32_t b_addr = a_addr + NACC * A_BYTES;
      const uint64_t db = make_smem_desc(b_addr);
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
        const uint64_t da = make_smem_desc(a_addr + a * A_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BK / TC_UK; ++k) {
          // advancing K inside the 128B swizzle row = +32 bytes on the start address (16 B units)
          umma_bf16(tmem_base + a * BMX, da + (uint64_t)(k * TC_UK * 2 >> 4), db + (uint64_t)(k * TC_UK * 2 >> 4), idesc,
                    (kb > 0 || k > 0) ? 1u : 0u);
        }
      }
      umma_commit(&empty_bar[s]);   // frees the smem stage once these MMAs have read it
    }
    umma_commit(accum_bar);         // accumulators complete
  }

  // ---------------- epilogue: all four warps ----------------
  __syncwarp();
  mbar_wait(accum_bar, 0);
  tc_fence_after();
  const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
  if (MODE == TC_GATE_BF16) {
    const int n = n0 + warp * 32 + lane;                     // this thread's gate row
#pragma unroll 1
    for (int c = 0; c < BMX; c += 32) {
      if (m0 + c >= p.M) break;                              // warp-uniform: whole chunk out of range
      float v[32], u[32];
      tmem_ld32(lane_addr + c, v);
      tmem_ld32(lane_addr + BMX + c, u);
      if (n < p.hidden) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int m = m0 + c + i;
          if (m < p.M) p.out_bf16[(long long)m * p.hidden + n] = __float2bfloat16_rn(silu(v[i]) * u[i]);
        }
      }
    }
  } else if (gridDim.z == 1) {
#pragma unroll 1
    for (int a = 0; a < NACC; ++a) {
      const int n = n0 + a * TC_BN + warp * 32 + lane;       // this thread's weight row in sub-tile a
#pragma unroll 1
      for (int c = 0; c < BMX; c += 32) {
        if (m0 + c >= p.M) break;
        float v[32];
        tmem_ld32(lane_addr + a * BMX + c, v);
        if (n < p.N) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int m = m0 + c + i;
            if (m < p.M) {
              float* dst = p.out_f32 + (long long)m * p.N + n;
              if (MODE == TC_STORE_F32) *dst = v[i];
              else *dst += v[i];
            }
          }
        }
      }
    }
  } else {
    // Split-K: this CTA's partial sums go to plane blockIdx.z of `partials` [S][M][N] with plain coalesced stores.
    // They are added IN SPLIT ORDER by the kernel that consumes the result anyway (RMSNorm for the residual
    // projections, RoPE + ring scatter for QKV, the greedy argmax for the text head; tc_reduce_partials for a
    // stand-alone GEMM), so the sum is the same in every run -- no atomics -- and costs no extra launch.
    float* plane = p.partials + (size_t)blockIdx.z * p.M * p.N;
#pragma unroll 1
    for (int a = 0; a < NACC; ++a) {
      const int n = n0 + a * TC_BN + warp * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BMX; c += 32) {
        if (m0 + c >= p.M) break;
        float v[32];
        tmem_ld32(lane_addr + a * BMX + c, v);
        if (n < p.N) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int m = m0 + c + i;
            if (m < p.M) plane[(long long)m * p.N + n] = v[i];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ================================================================================================
// host side
// ================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  DSM_CHECK(fn != nullptr, DSM_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  return fn;
}

// 2-D bf16 tensor [rows, cols] row-major; box = 64 columns (128 bytes) x box_rows, 128B swizzle.
static void make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {TC_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_tiled()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DSM_CHECK(r == CUDA_SUCCESS, DSM_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (rows=%llu cols=%llu box_rows=%u)", (int)r,
            (unsigned long long)rows, (unsigned long long)cols, box_rows);
}

bool tc_gemm_supported(int M, int N, int K, int mode) {
  if (M <= 0 || N <= 0 || K <= 0 || K % TC_BK != 0) return false;
  if (mode == TC_GATE_BF16 && (N % 2 != 0)) return false;
  return true;
}

// (Tried: half the stages, ~100 KB, so that two CTAs share an SM and one's prologue / epilogue overlaps the other's K
// loop, with twice the splits: LM branch 5.41 vs 5.26 ms -- slower, removed.)
template <int BMX, int MODE, int NACC>
struct TcCfg {
  static constexpr int STAGE_BYTES = NACC * TC_BN * TC_BK * 2 + BMX * TC_BK * 2;
  static constexpr int RAW = (200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = RAW > 8 ? 8 : RAW;
  static_assert(STAGES >= 2, "need at least two pipeline stages");
  static constexpr int SMEM = STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8 + 16 + 1024;   // barriers, tmem slot, last flag
};

template <int BMX, int MODE, int NACC>
static void launch_tc(const TcGemm& g, cudaStream_t st) {
  using C = TcCfg<BMX, MODE, NACC>;
  auto kern = tc_gemm_kernel<BMX, MODE, NACC, C::STAGES>;
  static PerDeviceOnce attr;
  attr.run([&] { DSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM)); });
  TcParams p;
  p.M = g.M, p.N = g.N, p.K = g.K, p.hidden = g.hidden;
  p.kb_total = g.K / TC_BK;
  p.kb_per_split = ceil_div(p.kb_total, g.split_k);
  DSM_CHECK(ceil_div(p.kb_total, p.kb_per_split) == g.split_k, DSM_ERR_INVALID_ARG, "tc_gemm: split_k %d leaves an empty split", g.split_k);
  p.out_f32 = g.out_f32, p.out_bf16 = g.out_bf16, p.partials = g.partials;
  const int n_rows = MODE == TC_GATE_BF16 ? g.hidden : g.N;
  dim3 grid(ceil_div(n_rows, TC_BN * (MODE == TC_GATE_BF16 ? 1 : NACC)), ceil_div(g.M, BMX), g.split_k);
  DSM_CHECK(g.split_k == 1 || (MODE != TC_GATE_BF16 && g.partials), DSM_ERR_INVALID_ARG, "tc_gemm: split_k %d without a partials buffer", g.split_k);
  launch_k(kern, grid, 128, C::SMEM, st, g.map_w, g.map_x, p);
  check_launch("tc_gemm");
}

void tc_gemm_prepare(TcGemm& g, const __nv_bfloat16* W, const __nv_bfloat16* X, int M, int N, int K, int mode,
                     float* out_f32, __nv_bfloat16* out_bf16, int sms_budget) {
  DSM_CHECK(tc_gemm_supported(M, N, K, mode), DSM_ERR_UNSUPPORTED, "tc_gemm: unsupported problem M=%d N=%d K=%d mode=%d", M, N,
            K, mode);
  g.M = M, g.N = N, g.K = K, g.mode = mode, g.out_f32 = out_f32, g.out_bf16 = out_bf16;
  g.hidden = mode == TC_GATE_BF16 ? N / 2 : 0;
  const int n_rows = mode == TC_GATE_BF16 ? g.hidden : N;
  // Two weight sub-tiles per CTA ("wide": 256 weight rows x 256 streams share one activation slab: 64 KB of operands
  // per k-block for twice the MACs of a 128 x 256 tile's 48 KB) for the large-batch regime, where the GEMMs are bound
  // by the L2 -> SM operand feed (profiles/r02l: ~5 TB/s).  Measured (LM branch alone, int8 ring, `tools/gpu_run_n.sh`):
  // at 2,000 streams it pays for out-proj / MLP-out, whose 128 narrow tiles would otherwise take an 8-way split with
  // 131 MB of partial planes (22.1 vs 23.2 ms); it is neutral for QKV / text head there, and at 1,024 streams it costs
  // 0.9 ms when it leaves QKV with 96 CTAs.  Hence: wide when the narrow grid is about a wave or more AND either the
  // wide grid still fills the SMs or the narrow one would not have.
  const long long sms = sms_budget;     // all SMs, or the LM partition of an ASR state (runtime.cuh)
  const long long narrow_tiles = (long long)ceil_div(n_rows, TC_BN) * ceil_div(M, 256);
  const long long wide_tiles = (long long)ceil_div(n_rows, 2 * TC_BN) * ceil_div(M, 256);
  const bool wide = mode != TC_GATE_BF16 && M > 256 && n_rows >= 2 * TC_BN && narrow_tiles >= 100 &&
                    (wide_tiles >= sms || narrow_tiles < sms);
  g.nacc = (mode == TC_GATE_BF16 || wide) ? 2 : 1;
  const int n_tiles = ceil_div(n_rows, TC_BN * (mode == TC_GATE_BF16 ? 1 : g.nacc));
  // Stream tile: as wide as the batch allows.  Every CTA must pull its W slab AND its X slab through
  // the SM's L2 port, so narrower stream tiles multiply the X traffic; to cover the 148 SMs we split K
  // instead (same total bytes, more CTAs); the partial sums are added in split order by the consuming kernel
  // (deterministic).  The gate pair has a non-linear epilogue, so it does not split.
  // The gate pair holds two accumulators: 128 streams per tile while that is what covers the SMs (256 streams: 88
  // CTAs instead of 44), 256 -- all 512 TMEM columns, twice the MMAs per operand byte -- once the batch fills them
  const long long gate_tiles_256 = (long long)ceil_div(n_rows, TC_BN) * ceil_div(M, 256);
  g.bmx = M <= 64 ? 64 : (M <= 128 ? 128 : ((mode == TC_GATE_BF16 && gate_tiles_256 < sms) ? 128 : 256));
  // DSM_TC_MAXSPLIT caps split-K (1 = no split: A/B of the partial-sum path)
  static const int env_split = [] { const char* e = getenv("DSM_TC_MAXSPLIT"); return e ? std::min(atoi(e), TC_MAX_SPLIT) : TC_MAX_SPLIT; }();
  g.split_k = 1;
  if (mode != TC_GATE_BF16) {
    // fill the 148 SMs by splitting K (at least 4 k-blocks per split) -- only when the tiles alone leave SMs idle:
    // every split writes and the consumer re-reads an [M, N] f32 plane, which at 2,000 streams is 50-250 MB per GEMM
    // (measured: 41.7 vs 37.1 ms per step with and without splits there)
    const int kb = K / TC_BK;
    const long long base = (long long)n_tiles * ceil_div(M, g.bmx);
    double best = 0.0;
    const long long wave = sms;
    // (measured at 2,000 streams: out-proj / MLP-out with 128 tiles and no split 140 us per launch, with 8 splits and
    // their 131 MB of partial planes 89 us -- the split stays whenever the tiles alone do not fill the SMs)
    for (int s = 1; s <= (base >= wave ? 1 : env_split); ++s) {
      const int per = ceil_div(kb, s);
      if (per < 4 && s > 1) break;
      const int s_eff = ceil_div(kb, per);
      const long long ctas = base * s_eff;
      const double eff = (double)ctas / (double)(ceil_div(ctas, wave) * wave);
      if (eff > best + 0.02) best = eff, g.split_k = s_eff;
    }
  }
  g.partial_bytes = g.split_k > 1 ? (size_t)g.split_k * M * N * sizeof(float) : 0;
  make_map(&g.map_w, W, (uint64_t)N, (uint64_t)K, TC_BN);
  make_map(&g.map_x, X, (uint64_t)M, (uint64_t)K, (uint32_t)g.bmx);
}

// out[m, n] (+)= sum_s partials[s][m][n], in split order: for GEMMs whose consumer does not do it (debug entry).
__global__ void tc_reduce_partials_kernel(const float* __restrict__ partials, int S, long long MN, int accumulate,
                                          float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < MN; i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += partials[(long long)s * MN + i];
    out[i] = accumulate ? out[i] + acc : acc;
  }
}
void tc_reduce_partials(const TcGemm& g, cudaStream_t st) {
  if (g.split_k <= 1) return;
  const long long MN = (long long)g.M * g.N;
  launch_k(tc_reduce_partials_kernel, (unsigned)std::min<long long>(ceil_div(MN, 256LL), 148 * 8), 256, 0, st, g.partials, g.split_k, MN,
           g.mode == TC_ACCUM_F32 ? 1 : 0, g.out_f32);
  check_launch("tc_reduce_partials");
}

void tc_gemm_launch(const TcGemm& g, cudaStream_t st) {
#define TC_CASE(BMX, MODE, NACC) \
  if (g.bmx == BMX && g.mode == MODE && g.nacc == NACC) return launch_tc<BMX, MODE, NACC>(g, st);
  TC_CASE(64, TC_STORE_F32, 1) TC_CASE(128, TC_STORE_F32, 1) TC_CASE(256, TC_STORE_F32, 1) TC_CASE(256, TC_STORE_F32, 2)
  TC_CASE(64, TC_ACCUM_F32, 1) TC_CASE(128, TC_ACCUM_F32, 1) TC_CASE(256, TC_ACCUM_F32, 1) TC_CASE(256, TC_ACCUM_F32, 2)
  TC_CASE(64, TC_GATE_BF16, 2) TC_CASE(128, TC_GATE_BF16, 2) TC_CASE(256, TC_GATE_BF16, 2)
#undef TC_CASE
  fail(DSM_ERR_UNSUPPORTED, "tc_gemm: no kernel for bmx=%d mode=%d nacc=%d", g.bmx, g.mode, g.nacc);
}

}  // namespace dsm
