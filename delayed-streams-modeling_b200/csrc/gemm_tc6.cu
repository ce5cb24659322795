// f32-accurate GEMM on the tcgen05 tensor pipe for the Mimi codec:  C[m, n] = sum_k A[m, k] * W[n, k],
// A and C in f32, by splitting every f32 operand into three bf16 terms (8 + 8 + 8 mantissa bits) and
// issuing the six significant cross products  a1w1 + a1w2 + a2w1 + a1w3 + a2w2 + a3w1  as kind::f16 MMAs
// that accumulate in f32 in TMEM.  The dropped terms are <= 2^-24 relative, so RVQ codes stay
// bit-identical to the f32 reference (TF32 or single-bf16 tiles flip ~1/3 of frames) while the work moves
// from the CUDA cores (~45 TFLOP/s achieved) to the tensor pipe (2.25 PFLOP/s / 6).
//
//   warp 0 lane 0 : TMA of the f32 A tile (implicit-GEMM conv window: a 3-D box over [B, T_out, k*C_in]
//                   with overlapping row stride) into a 128B-swizzled staging buffer
//   warp 3 lane 0 : TMA of the pre-split bf16 W planes (ring of stages)
//   warps 4-7     : converters -- read staging, split f32 -> 3 x bf16, write the three K-major 128B-
//                   swizzled UMMA operand tiles, fence.proxy.async, signal the MMA warp
//   warp 1 lane 0 : tcgen05.mma issuer: per k-block, for each of NJ column sub-tiles, 6 x 4 MMAs
//   warp 2        : TMEM alloc / dealloc
//   warps 8-15    : epilogue (tcgen05.ld -> bias / c2-dot / GELU / LayerScale / residual / ELU copy)
#include "gemm_tc6.cuh"

#include <mutex>

namespace dsm {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// One lane of a converged warp.  Unlike `lane == 0`, the compiler knows the guarded region runs in a single
// thread, so tcgen05.mma / TMA operands go straight to uniform registers; with `lane == 0` every tcgen05.mma was
// wrapped in a lane-election loop (~100 cycles per instruction, 3x the 32-48 cycles an N=64 MMA executes in).
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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
// A operand from tensor memory (lane = tile row, 32-bit column = two consecutive K elements), B from shared
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// Exact 3-way split of four floats by truncation (see the converter warps): x = h + m + l with bf16 h, m, l.
__device__ __forceinline__ void split3_x4(const float4& x, uint2& h, uint2& m, uint2& l) {
  const float xs[4] = {x.x, x.y, x.z, x.w};
  uint32_t hh[2], mm[2], ll[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const float a0 = xs[2 * e], a1 = xs[2 * e + 1];
    const uint32_t u0 = __float_as_uint(a0), u1 = __float_as_uint(a1);
    hh[e] = __byte_perm(u0, u1, 0x7632);
    const float r0 = a0 - __uint_as_float(u0 & 0xffff0000u), r1 = a1 - __uint_as_float(u1 & 0xffff0000u);
    const uint32_t v0 = __float_as_uint(r0), v1 = __float_as_uint(r1);
    mm[e] = __byte_perm(v0, v1, 0x7632);
    const float s0 = r0 - __uint_as_float(v0 & 0xffff0000u), s1 = r1 - __uint_as_float(v1 & 0xffff0000u);
    ll[e] = __byte_perm(__float_as_uint(s0), __float_as_uint(s1), 0x7632);
  }
  h = make_uint2(hh[0], hh[1]), m = make_uint2(mm[0], mm[1]), l = make_uint2(ll[0], ll[1]);
}
// ELU as the reference computes it: exp(x) - 1 for x <= 0 (candle Activation::Elu), not expm1
__device__ __forceinline__ float elu_ref(float x) { return x > 0.f ? x : expf(x) - 1.f; }
// K-major SWIZZLE_128B tile with 128-byte rows (see gemm_tc.cu)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

constexpr int T6_BM = 128;           // rows per CTA (UMMA M)
constexpr int T6_BK = 64;            // K elements per k-block
constexpr int T6_STG_SUB = T6_BM * 32 * 4;       // one staging sub-tile: 128 rows x 32 f32 (128-byte rows)
constexpr int T6_STG = 2 * T6_STG_SUB;           // 32 KB per k-block
constexpr int T6_PLANE = T6_BM * T6_BK * 2;      // 16 KB per bf16 plane tile
constexpr int T6_EPI_WARPS = 8;      // two per TMEM lane quadrant
constexpr int T6_CVT_WARPS = 4;      // one converter thread per tile row
constexpr int T6_THREADS = (4 + T6_CVT_WARPS + T6_EPI_WARPS) * 32;
// The f32-A variant keeps the converted A planes in TENSOR MEMORY (tcgen05.st by the converter warps, MMAs with A from
// TMEM); shared memory then holds three f32 staging slots and the W ring.  (The round-1 variant with the planes in
// shared memory measured slower -- codec step 2.98 vs 2.90 ms -- and is gone.)
constexpr int T6_NSTG = 3;                       // f32 staging slots
constexpr int T6_ACC_COLS_TS = 320;              // TMEM columns [0, 320): accumulators; [320, 512): two A buffers
constexpr int T6_A_COLS = 96;                    // TMEM columns of one A buffer: 3 planes x 32 (64 bf16 per row)
constexpr int T6_NPL_PRE = 3;        // operand-plane buffers of the pre-split variant (48 KB each)
constexpr int T6_SCR_LD = 36;        // floats per scratch row: 32 columns + 4 (float4 rows stay conflict-free)

}  // namespace

struct Tc6Params {
  int K, N;            // N = total output columns
  int T_out, B;        // logical rows = B * T_out, row m = b * T_out + t
  int t_sub, b_sub;    // rows of one tile = t_sub * b_sub (<= 128)
  int t_tiles;         // ceil(T_out / t_sub)
  int m_tiles;         // t_tiles * ceil(B / b_sub)
  int n_groups;        // ceil(N / (nj * BN))
  int nj;              // column sub-tiles per tile
  int pm;              // main-term accumulators the k-blocks rotate over (see the MMA issuer)
  int split_k;         // K splits per tile (work item = (tile, split)); partials meet in `ws`
  int acc_bufs;        // accumulator tile buffers in TMEM (1 or 2)
  float* ws;           // [tiles * split_k][128][nj * BN] partial sums (split_k > 1)
  int* counters;       // [tiles] arrivals; the last split to arrive reduces in split order and stores
  const float* bias;
  int neg;
  const float* scale;
  const float* res;
  RowMap r_map;
  int act;
  float* out_raw;
  RowMap o_map;
  float* out_act;
  RowMap e_map;
  __nv_bfloat16* out_pl;   // optional: the result as three bf16 planes (plane q at out_pl + q * pl_stride)
  RowMap pl_map;
  long long pl_stride;
};

// Persistent: gridDim.x CTAs (<= one per SM) walk the tiles; TMEM holds two accumulator buffers so the
// epilogue warps drain tile i while the producers / converters / MMA issuer are already on tile i+1.
// PRE: the A operand arrives already split (bf16 planes [3][rows][K] written by the producer of the activation:
// LayerNorm, attention, a GELU epilogue, the RVQ residual update), so the A tile goes from TMA straight into the
// swizzled operand buffers -- no f32 staging, no converter hop, and the freed 64 KB make the operand ring NPL deep.
// This is the form the codec's skinny GEMMs take (512 or 256 rows: one tile per CTA, latency-bound K loops).
template <int BN, int WS, bool PRE>
__global__ void __maxnreg__(128)   // 512 x 128 = the whole register file: one CTA per SM anyway (218 KB of shared memory)
tc6_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const Tc6Params p) {
  constexpr int W_STAGE = 3 * BN * T6_BK * 2;   // three bf16 planes of a BN x 64 tile
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // 128B-swizzle atoms need a 1024-byte aligned base
  constexpr bool TS = !PRE;                      // f32 A: planes in tensor memory
  constexpr int NPL = PRE ? T6_NPL_PRE : 2;      // operand-plane buffers (TS: in TMEM)
  constexpr int NSTG = T6_NSTG;
  uint8_t* stg = smem;                           // [NSTG][T6_STG]   (absent with PRE)
  uint8_t* planes = stg + (PRE ? 0 : NSTG * T6_STG);   // [NPL][3][T6_PLANE]   (absent with TS)
  uint8_t* wring = planes + (TS ? 0 : NPL * 3 * T6_PLANE);  // [WS][W_STAGE]
  constexpr int SR = PRE ? 8 : 16;               // rows of one epilogue warp's transposition tile
  constexpr int PARTS = 32 / SR, QN = SR / 4;
  float* scratch_all = reinterpret_cast<float*>(wring + WS * W_STAGE);   // [epilogue warps][SR][36]
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch_all + T6_EPI_WARPS * SR * T6_SCR_LD);
  uint64_t* stg_full = bars;             // [NSTG <= 4]
  uint64_t* stg_empty = bars + 4;        // [NSTG <= 4]
  uint64_t* pl_full = bars + 8;          // [NPL <= 4]
  uint64_t* pl_empty = bars + 12;        // [NPL <= 4]
  uint64_t* acc_full = bars + 16;        // [2]
  uint64_t* acc_empty = bars + 18;       // [2]
  uint64_t* w_full = bars + 20;          // [WS]
  uint64_t* w_empty = w_full + WS;       // [WS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + WS);
  uint32_t* last_flag = tmem_slot + 1;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;
  const int num_kb = (p.K + T6_BK - 1) / T6_BK;
  const int rows_box = p.t_sub * p.b_sub;
  const int acc_stride = p.nj * BN;          // columns of one accumulator set
  const int buf_stride = (1 + p.pm) * acc_stride;   // columns of one tile's accumulators
  const int AB = p.acc_bufs;                        // 2: the epilogue of tile i overlaps the MMAs of tile i+1
  const int total_tiles = p.m_tiles * p.n_groups;
  const int S = p.split_k;
  const int kb_per = (num_kb + S - 1) / S;       // the host picks S so that no split is empty
  const int total_items = total_tiles * S;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NPL; ++s) {
      mbar_init(&pl_full[s], PRE ? 1 : T6_CVT_WARPS);
      mbar_init(&pl_empty[s], 1);
    }
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(&stg_full[s], 1);
      mbar_init(&stg_empty[s], T6_CVT_WARPS);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], T6_EPI_WARPS);
    }
    for (int s = 0; s < WS; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // on-chip setup above; operands, residual and outputs belong to the kernels before this one

  if (warp == 0 && elect_one()) {
    // ---------------- A producer: f32 conv-window tile -> staging ----------------
    int kit = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int tile = item / S, kb0 = (item % S) * kb_per, kbn = min(kb_per, num_kb - kb0);
      const int mt = tile / p.n_groups;
      const int t0 = (mt % p.t_tiles) * p.t_sub, b0 = (mt / p.t_tiles) * p.b_sub;
      for (int kb = kb0; kb < kb0 + kbn; ++kb, ++kit) {
        if (PRE) {
          // three bf16 plane tiles of 128-byte rows straight into operand buffer kit % NPL
          const int s = kit % NPL;
          mbar_wait(&pl_empty[s], ((kit / NPL) & 1) ^ 1);
          mbar_expect_tx(&pl_full[s], (uint32_t)(3 * rows_box * 128));
          for (int q = 0; q < 3; ++q)
            tma_load_4d(planes + (s * 3 + q) * T6_PLANE, &map_a, kb * T6_BK, t0, b0, q, &pl_full[s]);
          continue;
        }
        const int s = kit % NSTG;
        mbar_wait(&stg_empty[s], ((kit / NSTG) & 1) ^ 1);
        const int kc = kb * T6_BK;
        const int nsub = (p.K - kc > 32) ? 2 : 1;
        mbar_expect_tx(&stg_full[s], (uint32_t)(nsub * rows_box * 128));
        for (int h = 0; h < nsub; ++h)
          tma_load_3d(stg + s * T6_STG + h * T6_STG_SUB, &map_a, kc + h * 32, t0, b0, &stg_full[s]);
      }
    }
  } else if (warp == 3 && elect_one()) {
    // ---------------- W producer: three bf16 planes per (k-block, column sub-tile) ----------------
    int wit = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int tile = item / S, kb0 = (item % S) * kb_per, kbn = min(kb_per, num_kb - kb0);
      const int n0 = (tile % p.n_groups) * p.nj * BN;
      const int nj = min(p.nj, (p.N - n0 + BN - 1) / BN);
      for (int kb = kb0; kb < kb0 + kbn; ++kb) {
        for (int j = 0; j < nj; ++j, ++wit) {
          const int s = wit % WS;
          mbar_wait(&w_empty[s], ((wit / WS) & 1) ^ 1);
          mbar_expect_tx(&w_full[s], (uint32_t)W_STAGE);
          for (int q = 0; q < 3; ++q)
            tma_load_3d(wring + s * W_STAGE + q * (BN * T6_BK * 2), &map_w, kb * T6_BK, n0 + j * BN, q, &w_full[s]);
        }
      }
    }
  } else if (warp == 1 && elect_one()) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = make_idesc(BN);
    // (a-plane, w-plane) pairs.  The tensor core truncates when it adds into the f32 accumulator, a
    // bias that grows with the number of accumulations, so the five small cross terms go to their
    // own accumulator (set 0) and the dominant a1*w1 term rotates over `pm` accumulators (sets 1..pm);
    // the epilogue adds the sets in registers with round-to-nearest.
    const int pa[6] = {0, 1, 2, 0, 1, 0};
    const int pw[6] = {2, 1, 0, 1, 0, 0};
    int kit = 0, wit = 0, li = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
      const int tile = item / S, kbn = min(kb_per, num_kb - (item % S) * kb_per);
      const int pm = min(p.pm, kbn);                       // main accumulators actually written
      const int n0 = (tile % p.n_groups) * p.nj * BN;
      const int nj = min(p.nj, (p.N - n0 + BN - 1) / BN);
      const int buf = li % AB;
      mbar_wait(&acc_empty[buf], ((li / AB) & 1) ^ 1);     // epilogue has drained this TMEM buffer
      tc_fence_after();
      const uint32_t acc_base = tmem_base + buf * buf_stride;
      for (int kb = 0; kb < kbn; ++kb, ++kit) {
        const int s = kit % NPL;
        mbar_wait(&pl_full[s], (kit / NPL) & 1);
        tc_fence_after();
        const uint32_t a_base = TS ? tmem_base + T6_ACC_COLS_TS + s * T6_A_COLS : smem_u32(planes + s * 3 * T6_PLANE);
        for (int j = 0; j < nj; ++j, ++wit) {
          const int ws = wit % WS;
          mbar_wait(&w_full[ws], (wit / WS) & 1);
          tc_fence_after();

          const uint32_t w_base = smem_u32(wring + ws * W_STAGE);
#pragma unroll
          for (int t = 0; t < 6; ++t) {
            const uint64_t da = make_smem_desc(a_base + pa[t] * T6_PLANE);
            const uint64_t dw = make_smem_desc(w_base + pw[t] * (BN * T6_BK * 2));
            const int set = t < 5 ? 0 : 1 + kb % pm;
            const bool first = t < 5 ? (kb == 0 && t == 0) : (kb < pm);
#pragma unroll
            for (int k = 0; k < T6_BK / 16; ++k) {
              if (TS)
                umma_bf16_ts(acc_base + set * acc_stride + j * BN, a_base + pa[t] * 32 + k * 8, dw + (uint64_t)(k * 2), idesc,
                             (first && k == 0) ? 0u : 1u);
              else
                umma_bf16(acc_base + set * acc_stride + j * BN, da + (uint64_t)(k * 2), dw + (uint64_t)(k * 2), idesc,
                          (first && k == 0) ? 0u : 1u);
            }
          }
          umma_commit(&w_empty[ws]);
        }
        umma_commit(&pl_empty[s]);
      }
      umma_commit(&acc_full[buf]);
    }
  } else if (!PRE && warp >= 4 && warp < 4 + T6_CVT_WARPS) {
    // ---------------- converters: staging f32 -> three swizzled bf16 operand tiles ----------------
    const int r = (warp - 4) * 32 + lane;    // tile row
    const int sw = r & 7;
    int kit = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int kb0 = (item % S) * kb_per, kbn = min(kb_per, num_kb - kb0);
      for (int kb = kb0; kb < kb0 + kbn; ++kb, ++kit) {
        const int s = kit % NSTG;
        const uint32_t ph = (kit / NSTG) & 1;
        const bool have2 = p.K - kb * T6_BK > 32;      // second 32-wide half present
        float4 v[16];
        mbar_wait(&stg_full[s], ph);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint8_t* src = stg + s * T6_STG + half * T6_STG_SUB + r * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            v[half * 8 + c] = (half == 0 || have2) ? *reinterpret_cast<const float4*>(src + ((c ^ sw) << 4))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // The slot goes back to the TMA producer, i.e. the async proxy will overwrite what the generic proxy just
        // read: that needs a proxy fence on this side as well, not only the mbarrier.  Without it whole tiles came
        // back wrong (about one k-block's worth, ~1e-2), non-deterministically and only under memory contention
        // (>= 64 CTAs) -- tools/gemm6_stress.py; an A/B build with only this fence added is clean.
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&stg_empty[s]);      // staging slot may be refilled
        const int ps = kit % NPL;
        mbar_wait(&pl_empty[ps], ((kit / NPL) & 1) ^ 1);  // MMAs that read this plane buffer are done
        tc_fence_after();
        // TS: thread r owns TMEM lane r (the converter warps 4..7 reach lane quadrants 0..3); 32 columns per plane
        const uint32_t tdst = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + T6_ACC_COLS_TS + ps * T6_A_COLS;
        uint32_t h1[4], h2[4], h3[4];
        // Exact 3-way split by TRUNCATION: b1 = top 16 bits of x, r1 = x - b1 (exact), b2 = top 16 bits of
        // r1, r2 = r1 - b2 (exact), b3 = top 16 bits of r2 (r2 has <= 8 significant bits left, so b3 = r2).
        // No F2F conversions: one PRMT packs the upper halves of two floats into a bf16x2.
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const float4 x0 = v[2 * cc], x1 = v[2 * cc + 1];
          const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
          uint32_t p1[4], p2[4], p3[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a0 = xs[2 * e], a1 = xs[2 * e + 1];
            const uint32_t u0 = __float_as_uint(a0), u1 = __float_as_uint(a1);
            p1[e] = __byte_perm(u0, u1, 0x7632);
            const float r0 = a0 - __uint_as_float(u0 & 0xffff0000u), r1 = a1 - __uint_as_float(u1 & 0xffff0000u);
            const uint32_t v0 = __float_as_uint(r0), v1 = __float_as_uint(r1);
            p2[e] = __byte_perm(v0, v1, 0x7632);
            const float s0 = r0 - __uint_as_float(v0 & 0xffff0000u), s1 = r1 - __uint_as_float(v1 & 0xffff0000u);
            p3[e] = __byte_perm(__float_as_uint(s0), __float_as_uint(s1), 0x7632);
          }
          {
            // eight floats -> four packed columns per plane; stored eight columns at a time
            if (cc & 1) {
              uint32_t q1[8] = {h1[0], h1[1], h1[2], h1[3], p1[0], p1[1], p1[2], p1[3]};
              uint32_t q2[8] = {h2[0], h2[1], h2[2], h2[3], p2[0], p2[1], p2[2], p2[3]};
              uint32_t q3[8] = {h3[0], h3[1], h3[2], h3[3], p3[0], p3[1], p3[2], p3[3]};
              tmem_st8(tdst + (cc - 1) * 4, q1);
              tmem_st8(tdst + 32 + (cc - 1) * 4, q2);
              tmem_st8(tdst + 64 + (cc - 1) * 4, q3);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) h1[e] = p1[e], h2[e] = p2[e], h3[e] = p3[e];
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pl_full[ps]);
      }
    }
  } else if (warp >= 4 + T6_CVT_WARPS) {
    // ---------------- epilogue warps: two per TMEM lane quadrant ----------------
    // A thread owns one TMEM lane = one output row, so a direct store would scatter 16-byte pieces over
    // 32 rows per instruction; each warp transposes its 32x32 chunk through shared memory so that
    // consecutive lanes write consecutive columns of one row (and reads the residual the same way).
    // The two warps of a quadrant take alternate chunks; with a single chunk they split its rows.
    const int ew = warp - (4 + T6_CVT_WARPS);
    const int quad = warp & 3, sib = ew >> 2;     // sibling index 0/1 within the quadrant
    float* scratch = scratch_all + ew * SR * T6_SCR_LD;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    // (Tried: the two sibling groups taking alternate tiles, one TMEM buffer each, so that two tiles' epilogues are in
    // flight per SM -- measured slower, codec step 3.10 vs 2.87 ms: a tile then drains through 4 warps instead of 8.)
    int li = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
      const int tile = item / S, split = item % S;
      const int pm = min(p.pm, min(kb_per, num_kb - split * kb_per));
      const int tile_n = p.nj * BN;
      float* ws_mine = p.ws + ((size_t)(tile * S + split) * T6_BM + quad * 32) * tile_n;
      const int mt = tile / p.n_groups;
      const int t0 = (mt % p.t_tiles) * p.t_sub, b0 = (mt / p.t_tiles) * p.b_sub;
      const int n0 = (tile % p.n_groups) * p.nj * BN;
      const int nj = min(p.nj, (p.N - n0 + BN - 1) / BN);
      const int r = quad * 32 + lane;
      const int bl = r / p.t_sub, tl = r - bl * p.t_sub;
      const int b = b0 + bl, t = t0 + tl;
      const uint32_t row_ok = (r < rows_box && b < p.B && t < p.T_out) ? 1u : 0u;
      const int m = row_ok ? b * p.T_out + t : 0;
      const uint32_t res_off = p.res ? (uint32_t)p.r_map.off(m) : 0u;
      const uint32_t raw_off = p.out_raw ? (uint32_t)p.o_map.off(m) : 0u;
      const uint32_t act_off = p.out_act ? (uint32_t)p.e_map.off(m) : 0u;
      const uint32_t pl_off = p.out_pl ? (uint32_t)p.pl_map.off(m) : 0u;
      const uint32_t ok_mask = __ballot_sync(0xffffffffu, row_ok);
      const int buf = li % AB;
      mbar_wait(&acc_full[buf], (li / AB) & 1);
      tc_fence_after();
      const uint32_t acc_addr = lane_addr + buf * buf_stride;
      const int n_chunks = nj * BN / 32;
      const bool split_rows = n_chunks == 1;       // both siblings work on chunk 0, half the rows each
      const int c_first = split_rows ? 0 : sib, c_step = split_rows ? 1 : 2;
      const int qrow = lane >> 3, cs = lane & 7;   // 8 lanes per row: 8 x float4 = one 128-byte segment
#pragma unroll 1
      for (int c = c_first; c < n_chunks; c += c_step) {
        // The residual rows of this chunk are requested first (8 float4 per thread), so that their HBM round trip
        // overlaps the TMEM loads and the transposition instead of following them.
        const int col = n0 + c * 32 + cs * 4;
        const bool col_ok = col < p.N;              // N is a multiple of 32: a chunk is valid as a whole
        float4 rres[PARTS][QN];
#pragma unroll
        for (int half = 0; half < PARTS; ++half) {
#pragma unroll
          for (int q = 0; q < QN; ++q) {
            const int rr = half * SR + 4 * q + qrow;
            const uint32_t ro = __shfl_sync(0xffffffffu, res_off, rr);
            const bool want = p.res && col_ok && ((ok_mask >> rr) & 1u) && !(split_rows && half / (PARTS / 2) != sib) && S == 1;
            // `res` may alias `out_raw` (x <- x + ...): a coherent load, not ld.global.nc
            rres[half][q] = want ? __ldcg(reinterpret_cast<const float4*>(p.res + ro + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        float v[32];
        tmem_ld32(acc_addr + c * 32, v);
        for (int a = 1; a <= pm; ++a) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float u[16];
            tmem_ld16(acc_addr + a * acc_stride + c * 32 + h * 16, u);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[h * 16 + i] += u[i];
          }
        }
        // Transposition through shared memory, SR (16 or 8) rows x 32 columns at a time: the lanes that own the
        // SR TMEM rows park them as float4s in a [SR][36] scratch tile, then all 32 lanes sweep it four
        // rows per instruction (8 lanes x 16 B = one full 128-byte row segment), so every global access
        // of the epilogue is a float4 in a fully used line.
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), scale4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (p.bias && col_ok) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        if (p.scale && col_ok) scale4 = __ldg(reinterpret_cast<const float4*>(p.scale + col));
#pragma unroll
        for (int half = 0; half < PARTS; ++half) {
          if (split_rows && half / (PARTS / 2) != sib) continue;  // warp-uniform
          if (lane / SR == half) {
            float4* dst = reinterpret_cast<float4*>(scratch + (lane % SR) * T6_SCR_LD);
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          __syncwarp();
          if (S > 1) {
            // split-K: park the raw partial sums of this (tile, split) in the workspace
#pragma unroll
            for (int q = 0; q < QN; ++q) {
              const int rr = half * SR + 4 * q + qrow;
              *reinterpret_cast<float4*>(ws_mine + (size_t)rr * tile_n + c * 32 + cs * 4) =
                  *reinterpret_cast<const float4*>(scratch + (4 * q + qrow) * T6_SCR_LD + cs * 4);
            }
            __syncwarp();
            continue;
          }
#pragma unroll
          for (int q = 0; q < QN; ++q) {
            const int rr = half * SR + 4 * q + qrow;
            const uint32_t wo = __shfl_sync(0xffffffffu, raw_off, rr);
            const uint32_t ao = __shfl_sync(0xffffffffu, act_off, rr);
            const uint32_t po = __shfl_sync(0xffffffffu, pl_off, rr);
            if (!(col_ok && ((ok_mask >> rr) & 1u))) continue;
            float4 x = *reinterpret_cast<const float4*>(scratch + (4 * q + qrow) * T6_SCR_LD + cs * 4);
            if (p.bias) {
              if (p.neg) x.x = bias4.x - x.x, x.y = bias4.y - x.y, x.z = bias4.z - x.z, x.w = bias4.w - x.w;
              else x.x += bias4.x, x.y += bias4.y, x.z += bias4.z, x.w += bias4.w;
            }
            if (p.act == 1) x.x = gelu_erf(x.x), x.y = gelu_erf(x.y), x.z = gelu_erf(x.z), x.w = gelu_erf(x.w);
            x.x = x.x * scale4.x + rres[half][q].x, x.y = x.y * scale4.y + rres[half][q].y;
            x.z = x.z * scale4.z + rres[half][q].z, x.w = x.w * scale4.w + rres[half][q].w;
            if (p.out_raw) *reinterpret_cast<float4*>(p.out_raw + wo + col) = x;
            if (p.out_act)
              *reinterpret_cast<float4*>(p.out_act + ao + col) =
                  make_float4(elu_ref(x.x), elu_ref(x.y), elu_ref(x.z), elu_ref(x.w));
            if (p.out_pl) {
              // the consumer GEMM's A operand, already split into three bf16 planes
              uint2 h, md, lo;
              split3_x4(x, h, md, lo);
              *reinterpret_cast<uint2*>(p.out_pl + po + col) = h;
              *reinterpret_cast<uint2*>(p.out_pl + p.pl_stride + po + col) = md;
              *reinterpret_cast<uint2*>(p.out_pl + 2 * p.pl_stride + po + col) = lo;
            }
          }
          __syncwarp();
        }
      }
      // this warp is done reading the TMEM buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (S > 1) {
        // The split that arrives last for this tile adds all partials IN SPLIT ORDER (deterministic,
        // unlike atomics) and runs the epilogue.
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(T6_EPI_WARPS * 32) : "memory");
        if (ew == 0 && lane == 0) {
          const int old = atomicAdd(p.counters + tile, 1);
          const uint32_t last = old == S - 1 ? 1u : 0u;
          if (last) p.counters[tile] = 0;      // ready for the next launch (graph replay)
          *last_flag = last;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(T6_EPI_WARPS * 32) : "memory");
        const bool is_last = *last_flag != 0;
        if (is_last) {
          // All eight epilogue warps sweep the tile linearly, four columns per thread, every split's
          // partial in flight at once; each element is summed in split order.
          __threadfence();
          const float* ws_tile = p.ws + (size_t)tile * S * T6_BM * tile_n;
          const size_t split_stride = (size_t)T6_BM * tile_n;
          const int et = ew * 32 + lane;
          const int tile_cols = min(tile_n, p.N - n0);
          for (int e = et * 4; e < T6_BM * tile_n; e += T6_EPI_WARPS * 32 * 4) {
            const int r = e / tile_n, cc = e - r * tile_n;
            const int bl = r / p.t_sub, tl = r - bl * p.t_sub;
            const int b = b0 + bl, t = t0 + tl;
            if (r >= rows_box || b >= p.B || t >= p.T_out || cc >= tile_cols) continue;
            const int m = b * p.T_out + t;
            const int col = n0 + cc;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int sp = 0; sp < S; ++sp) {
              const float4 v = __ldcg(reinterpret_cast<const float4*>(ws_tile + sp * split_stride + e));
              x.x += v.x, x.y += v.y, x.z += v.z, x.w += v.w;
            }
            float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float y = xs[i];
              if (p.bias) y = p.neg ? __ldg(p.bias + col + i) - y : y + __ldg(p.bias + col + i);
              if (p.act == 1) y = gelu_erf(y);
              if (p.scale) y *= __ldg(p.scale + col + i);
              xs[i] = y;
            }
            if (p.res) {
              const float4 q = *reinterpret_cast<const float4*>(p.res + p.r_map.off(m) + col);
              xs[0] += q.x, xs[1] += q.y, xs[2] += q.z, xs[3] += q.w;
            }
            if (p.out_raw)
              *reinterpret_cast<float4*>(p.out_raw + p.o_map.off(m) + col) = make_float4(xs[0], xs[1], xs[2], xs[3]);
            if (p.out_act)
              *reinterpret_cast<float4*>(p.out_act + p.e_map.off(m) + col) =
                  make_float4(elu_ref(xs[0]), elu_ref(xs[1]), elu_ref(xs[2]), elu_ref(xs[3]));
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(T6_EPI_WARPS * 32) : "memory");   // last_flag may be rewritten
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ================================================================================================
// host side
// ================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled6() {
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

template <int BN, bool PRE>
struct T6Cfg {
  static constexpr int W_STAGE = 3 * BN * T6_BK * 2;
  // f32 staging (2 x 32 KB) + two operand-plane buffers, or three plane buffers fed directly by TMA
  static constexpr int FIXED = PRE ? T6_NPL_PRE * 3 * T6_PLANE : T6_NSTG * T6_STG;
  static constexpr int SCRATCH = T6_EPI_WARPS * (PRE ? 8 : 16) * T6_SCR_LD * 4;
  static constexpr int WS_RAW = (232448 - 512 - FIXED - SCRATCH) / W_STAGE;
  static constexpr int WS = WS_RAW > 6 ? 6 : WS_RAW;
  static constexpr int SMEM = FIXED + WS * W_STAGE + SCRATCH + (20 + 2 * WS) * 8 + 16;   // barriers + tmem slot + last flag
  static_assert(SMEM <= 232448, "exceeds the 227 KB shared-memory limit per CTA");
  static_assert(WS >= 2, "need at least two W stages");
};

template <int BN, bool PRE>
static void launch6(const Tc6Gemm& g, cudaStream_t st) {
  using C = T6Cfg<BN, PRE>;
  auto kern = tc6_gemm_kernel<BN, C::WS, PRE>;
  static PerDeviceOnce attr;
  attr.run([&] { DSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM)); });
  Tc6Params p;
  p.K = g.K, p.N = g.N, p.T_out = g.T_out, p.B = g.B, p.t_sub = g.t_sub, p.b_sub = g.b_sub;
  p.t_tiles = ceil_div(g.T_out, g.t_sub), p.nj = g.nj, p.pm = g.pm;
  p.m_tiles = p.t_tiles * ceil_div(g.B, g.b_sub);
  p.n_groups = ceil_div(g.N, g.nj * BN);
  p.split_k = g.split_k, p.ws = g.ws, p.counters = g.counters;
  p.acc_bufs = g.acc_bufs;
  p.bias = g.ep.bias, p.neg = g.ep.neg, p.scale = g.ep.scale, p.res = g.ep.res, p.r_map = g.ep.r_map, p.act = g.ep.act;
  p.out_raw = g.ep.out_raw, p.o_map = g.ep.o_map, p.out_act = g.ep.out_act, p.e_map = g.ep.e_map;
  p.out_pl = g.ep.out_planes, p.pl_map = g.ep.p_map, p.pl_stride = g.ep.out_plane_stride;
  const long long tiles = (long long)p.m_tiles * p.n_groups * g.split_k;
  const int grid = (int)std::min<long long>(tiles, g.sms);
  launch_k(kern, grid, T6_THREADS, C::SMEM, st, g.map_a, g.map_w, p);
  check_launch("tc6_gemm");
}

bool tc6_supported(const GemmParams& g) {
  if (g.M <= 0 || g.N % 32 != 0 || g.K % 32 != 0 || g.out_bf16) return false;
  if (g.a_map.T <= 0 || g.M % g.a_map.T != 0) return false;
  // TMA: strides in bytes must be multiples of 16
  const int esz = g.a_planes ? 2 : 4;
  if ((g.a_map.stride_t * esz) % 16 != 0 || (g.a_map.stride_b * esz) % 16 != 0) return false;
  if (g.a_planes && (g.a_plane_stride * 2) % 16 != 0) return false;
  if (reinterpret_cast<uintptr_t>(g.a_planes ? (const void*)g.a_planes : g.A) % 16 != 0) return false;
  if (g.out_planes && (reinterpret_cast<uintptr_t>(g.out_planes) % 8 != 0 || g.p_map.stride_t % 4 != 0 ||
                       g.p_map.stride_b % 4 != 0 || g.out_plane_stride % 4 != 0))
    return false;
  // the epilogue reads and writes float4s: every row start must be 16-byte aligned
  auto vec_ok = [](const void* ptr, const RowMap& m) {
    return ptr == nullptr || (reinterpret_cast<uintptr_t>(ptr) % 16 == 0 && m.stride_t % 4 == 0 && m.stride_b % 4 == 0);
  };
  if (!vec_ok(g.res, g.r_map) || !vec_ok(g.out_raw, g.o_map) || !vec_ok(g.out_act, g.e_map)) return false;
  if (reinterpret_cast<uintptr_t>(g.bias) % 16 != 0 || reinterpret_cast<uintptr_t>(g.scale) % 16 != 0) return false;
  return true;
}

__nv_bfloat16* tc6_split_weights(Arena& arena, const std::vector<float>& w, int N, int K) {
  DSM_CHECK((size_t)N * K == w.size(), DSM_ERR_SHAPE, "tc6_split_weights: %d x %d != %zu", N, K, w.size());
  std::vector<uint16_t> planes((size_t)3 * N * K);
  auto to_bf16 = [](float x) -> uint16_t {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
  };
  auto from_bf16 = [](uint16_t h) -> float {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
  };
  const size_t n = (size_t)N * K;
  for (size_t i = 0; i < n; ++i) {
    const float x = w[i];
    const uint16_t b1 = to_bf16(x);
    const float r1 = x - from_bf16(b1);
    const uint16_t b2 = to_bf16(r1);
    const float r2 = r1 - from_bf16(b2);
    planes[i] = b1, planes[n + i] = b2, planes[2 * n + i] = to_bf16(r2);
  }
  __nv_bfloat16* d = arena.alloc<__nv_bfloat16>(3 * n, false);
  DSM_CUDA(cudaMemcpy(d, planes.data(), 3 * n * 2, cudaMemcpyHostToDevice));
  return d;
}

void tc6_prepare(Tc6Gemm& t, const GemmParams& g, const __nv_bfloat16* w_planes, const Tc6Workspace& wsp, int sms) {
  DSM_CHECK(tc6_supported(g), DSM_ERR_UNSUPPORTED, "tc6_gemm: unsupported problem M=%d N=%d K=%d", g.M, g.N, g.K);
  t.ep = g;
  t.K = g.K, t.N = g.N, t.T_out = g.a_map.T, t.B = g.M / g.a_map.T;
  // the epilogue keeps row offsets in 32 bits
  for (const RowMap* rm : {&g.r_map, &g.o_map, &g.e_map}) {
    const long long extent = rm->off(g.M - 1) + g.N;
    DSM_CHECK(extent < (1LL << 32), DSM_ERR_UNSUPPORTED, "tc6_gemm: output extent %lld exceeds 32-bit element offsets", extent);
  }
  // rows of a tile: a run of time steps of one stream, or whole (short) streams side by side
  if (t.T_out >= T6_BM || T6_BM % t.T_out != 0) t.t_sub = std::min(t.T_out, T6_BM), t.b_sub = 1;
  else t.t_sub = t.T_out, t.b_sub = T6_BM / t.T_out;
  t.bn = g.N % 64 == 0 ? 64 : 32;
  const long long m_tiles = (long long)ceil_div(t.T_out, t.t_sub) * ceil_div(t.B, t.b_sub);
  const int n_sub = ceil_div(g.N, t.bn);
  // column sub-tiles per CTA: as many as TMEM holds (512 columns) while the grid still covers the SMs
  // (TMEM: two tile buffers of 256 columns, each shared by the correction accumulator and >= 1 main
  // accumulator => at most 128 output columns per tile)
  constexpr int max_tile_n = 64;
  // pre-split A: one column sub-tile per CTA keeps the W ring as deep (in k-blocks) as the operand ring
  int nj = std::max(1, std::min(n_sub, std::min(g.a_planes ? 64 : 128, max_tile_n) / t.bn));
  t.sms = sms;
  while (nj > 1 && m_tiles * ceil_div(n_sub, nj) < sms) nj = (nj + 1) / 2;
  t.nj = nj;
  // TMEM: 512 accumulator columns, or 320 when the A planes live in tensor memory (f32-A variant).  Two tile
  // buffers when they fit with at least one main accumulator beside the correction accumulator; long K loops
  // (> 8 k-blocks) want more rotating accumulators than that leaves and take the whole region as one buffer
  // (the tensor core truncates when it accumulates: see the MMA issuer).
  {
    const bool ts = !g.a_planes;
    const int acc_cols = ts ? T6_ACC_COLS_TS : 512;
    const int tile_n = nj * t.bn;
    const bool long_k = ts && ceil_div(g.K, T6_BK) > 8;
    t.acc_bufs = (!long_k && 2 * 2 * tile_n <= acc_cols) ? 2 : 1;
    t.pm = std::max(1, std::min(7, acc_cols / (t.acc_bufs * tile_n) - 1));
  }
  // Skinny problems (few tiles, long K) cannot fill 148 SMs with tiles alone: split K over several CTAs
  // per tile; the partial sums meet in the shared workspace and the last arrival reduces them in order.
  {
    const long long tiles = m_tiles * ceil_div(n_sub, nj);
    const int num_kb = ceil_div(g.K, T6_BK);
    int S = 1;
    // On by default since round 2 (measured again on the final kernel: codec step 2.79 vs 2.92 ms at 256 streams, 1.69
    // vs 1.85 ms at 64, no change at 2,000; round 1 had measured a small loss on a slower kernel).  DSM_TC6_SPLITK=0
    // turns it off; only considered where >= 6 k-blocks remain per split.
    static const bool splitk_on = [] { const char* e = getenv("DSM_TC6_SPLITK"); return !(e && e[0] == '0'); }();
    // (the last-arrival reduce path does not write plane outputs: such GEMMs never split)
    if (splitk_on && wsp.ws && tiles <= sms / 2 && num_kb >= 16 && !g.out_planes) {
      S = (int)std::max<long long>(1, std::min<long long>(sms / tiles, num_kb / 6));
      const int per = ceil_div(num_kb, S);
      S = ceil_div(num_kb, per);
      const size_t need = (size_t)tiles * S * T6_BM * nj * t.bn * sizeof(float);
      if (need > wsp.ws_bytes || tiles > wsp.n_counters) S = 1;
    }
    t.split_k = S, t.ws = wsp.ws, t.counters = wsp.counters;
  }
  t.presplit = g.a_planes != nullptr;
  if (t.presplit) {
    // A: bf16 planes [3][B][T_out][K] window view, box 64 x t_sub x b_sub x 1 (128-byte rows, 128B swizzle: the
    // tile lands in the UMMA operand layout)
    cuuint64_t gdim[4] = {(cuuint64_t)g.K, (cuuint64_t)t.T_out, (cuuint64_t)t.B, 3};
    cuuint64_t gstr[3] = {(cuuint64_t)g.a_map.stride_t * 2, (cuuint64_t)g.a_map.stride_b * 2, (cuuint64_t)g.a_plane_stride * 2};
    if (t.T_out == 1) gstr[0] = gstr[1];
    cuuint32_t box[4] = {64, (cuuint32_t)t.t_sub, (cuuint32_t)t.b_sub, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode_tiled6()(&t.map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(g.a_planes), gdim,
                                 gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DSM_CHECK(r == CUDA_SUCCESS, DSM_ERR_CUDA, "tc6 pre-split A tensor map failed (%d): K=%d T=%d B=%d", (int)r, g.K, t.T_out, t.B);
  } else {
    cuuint64_t gdim[3] = {(cuuint64_t)g.K, (cuuint64_t)t.T_out, (cuuint64_t)t.B};
    cuuint64_t gstr[2] = {(cuuint64_t)g.a_map.stride_t * 4, (cuuint64_t)g.a_map.stride_b * 4};
    if (t.T_out == 1) gstr[0] = gstr[1];      // degenerate time dimension: any legal stride
    cuuint32_t box[3] = {32, (cuuint32_t)t.t_sub, (cuuint32_t)t.b_sub};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode_tiled6()(&t.map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(g.A), gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DSM_CHECK(r == CUDA_SUCCESS, DSM_ERR_CUDA, "tc6 A tensor map failed (%d): K=%d T=%d B=%d strides %lld %lld", (int)r, g.K,
              t.T_out, t.B, (long long)gstr[0], (long long)gstr[1]);
  }
  // W: bf16 [3][N][K] planes, box 64 x bn x 1
  {
    cuuint64_t gdim[3] = {(cuuint64_t)g.K, (cuuint64_t)g.N, 3};
    cuuint64_t gstr[2] = {(cuuint64_t)g.K * 2, (cuuint64_t)g.K * g.N * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)t.bn, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode_tiled6()(&t.map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(w_planes), gdim, gstr,
                                 box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DSM_CHECK(r == CUDA_SUCCESS, DSM_ERR_CUDA, "tc6 W tensor map failed (%d): N=%d K=%d", (int)r, g.N, g.K);
  }
}

Tc6Workspace tc6_make_workspace(Arena& arena) {
  Tc6Workspace w;
  w.ws_bytes = (size_t)16 << 20;     // 148 work items x 128 rows x 128 columns x 4 B = 9.7 MB at most
  w.ws = arena.alloc<float>(w.ws_bytes / sizeof(float));
  w.n_counters = 1024;
  w.counters = arena.alloc<int>(w.n_counters);
  return w;
}

void tc6_launch(const Tc6Gemm& t, cudaStream_t st) {
  if (t.presplit) return t.bn == 64 ? launch6<64, true>(t, st) : launch6<32, true>(t, st);
  return t.bn == 64 ? launch6<64, false>(t, st) : launch6<32, false>(t, st);
}

}  // namespace dsm
