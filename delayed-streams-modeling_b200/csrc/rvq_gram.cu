// Grouped residual-VQ encode: the argmin chain of G consecutive codebooks as ONE kernel behind ONE score GEMM.
//
// The reference quantises codebook by codebook (quantization.rs:219-229): codes_q = argmin_n (|e_n|^2/2 - r_q . e_n),
// r_{q+1} = r_q - e_q[codes_q] -- 32 dependent (GEMM, select) pairs of ~21 us each at 256 streams, a quarter of the
// codec step and all launch latency.  Within a group q0 .. q0+G-1 the residual is r_q = r_q0 - sum_{p<q} e_p[codes_p], so
//
//     |e_n|^2/2 - r_q . e_n  =  (|e_n|^2/2 - r_q0 . e_n)  +  sum_{q0 <= p < q} (e_p[codes_p] . e_n)
//
// The first term for ALL G codebooks of the group is one [B, qdim] x [qdim, G * bins] GEMM on the group's entry
// residual (the tcgen05 codec GEMM, same kernel and operand split as the per-codebook scores); the second is a row of
// the constant Gram table E_p E_q^T, built once at load in double precision and rounded to f32.  What is left of the
// dependency chain is one CTA per stream that, per codebook, adds the rows its earlier choices select (8 KB each,
// independent loads) and takes the first-index argmin: ~2 us per codebook instead of 21.  The residual is then
// updated exactly as the reference does (one f32 subtraction per codebook, in codebook order), so the next group's
// GEMM sees bit-identical input; the regrouping only changes how r_q . e_n is rounded, and the group size bounds the
// cancellation (G = 1 is the reference association).  Tables: G (G - 1) / 2 pairs x bins^2 f32 per group -- 1.8 GB
// for 31 codebooks in groups of 8.
#include "kernels.cuh"

namespace dsm {

// G[i][j] = sum_d Ep[i][d] * Eq[j][d], accumulated in double (each product is exact in double), rounded once.
__global__ void __launch_bounds__(256) rvq_gram_table_kernel(const float* __restrict__ Ep, const float* __restrict__ Eq, int bins,
                                                              int qd, float* __restrict__ G) {
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  double acc[2][2] = {{0, 0}, {0, 0}};
  for (int d0 = 0; d0 < qd; d0 += 32) {
    for (int t = threadIdx.x; t < 32 * 32; t += 256) {
      const int r = t / 32, c = t % 32;
      sa[r][c] = (i0 + r < bins && d0 + c < qd) ? Ep[(long long)(i0 + r) * qd + d0 + c] : 0.f;
      sb[r][c] = (j0 + r < bins && d0 + c < qd) ? Eq[(long long)(j0 + r) * qd + d0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < 32; ++c) {
      const double a0 = sa[ty][c], a1 = sa[ty + 16][c], b0 = sb[tx][c], b1 = sb[tx + 16][c];
      acc[0][0] = fma(a0, b0, acc[0][0]), acc[0][1] = fma(a0, b1, acc[0][1]);
      acc[1][0] = fma(a1, b0, acc[1][0]), acc[1][1] = fma(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
      if (i < bins && j < bins) G[(long long)i * bins + j] = (float)acc[a][b];
    }
}

void launch_rvq_gram_table(const float* Ep, const float* Eq, int bins, int qd, float* G, cudaStream_t st) {
  rvq_gram_table_kernel<<<dim3((bins + 31) / 32, (bins + 31) / 32), 256, 0, st>>>(Ep, Eq, bins, qd, G);
  check_launch("rvq_gram_table");
}

// One CTA per stream.  scores [B][g][bins] = |e|^2/2 - r_q0 . e for the group's codebooks; gram = the group's pair
// tables, pair (p < q) at index p*g - p(p+1)/2 + (q - p - 1), each [bins][bins]; emb = the group's codebooks
// [g][bins][qd].  Ties resolve to the first index like candle's argmin (quantization.rs:131).
template <int MAXG, int NV>
__global__ void __launch_bounds__(256) rvq_chain_kernel(const float* __restrict__ scores, int g, int bins,
                                                         const float* __restrict__ gram, const float* __restrict__ emb, int qd,
                                                         float* __restrict__ residual, long long res_stride,
                                                         uint32_t* __restrict__ codes, int n_q, int q0,
                                                         __nv_bfloat16* __restrict__ planes, long long plane_stride) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ int s_code[MAXG];
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* row = scores + (long long)b * g * bins;
  const long long tbl = (long long)bins * bins;
  for (int q = 0; q < g; ++q) {
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = (i * 256 + tid) * 4;
      v[i] = j < bins ? *reinterpret_cast<const float4*>(row + (long long)q * bins + j) : make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
    }
#pragma unroll
    for (int p = 0; p < MAXG - 1; ++p) {
      if (p < q) {
        const float* t = gram + (long long)(p * g - p * (p + 1) / 2 + (q - p - 1)) * tbl + (long long)s_code[p] * bins;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int j = (i * 256 + tid) * 4;
          if (j < bins) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(t + j));
            v[i].x += a.x, v[i].y += a.y, v[i].z += a.z, v[i].w += a.w;
          }
        }
      }
    }
    float best = INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = (i * 256 + tid) * 4;                 // ascending per thread: strict < keeps the first
      if (v[i].x < best) best = v[i].x, bi = j;
      if (v[i].y < best) best = v[i].y, bi = j + 1;
      if (v[i].z < best) best = v[i].z, bi = j + 2;
      if (v[i].w < best) best = v[i].w, bi = j + 3;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov < best || (ov == best && oi < bi)) best = ov, bi = oi;
    }
    if ((tid & 31) == 0) s_val[tid >> 5] = best, s_idx[tid >> 5] = bi;
    __syncthreads();
    best = s_val[0], bi = s_idx[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float ov = s_val[w];
      const int oi = s_idx[w];
      if (ov < best || (ov == best && oi < bi)) best = ov, bi = oi;
    }
    if (bi < 0 || bi >= bins) bi = 0;                    // all-NaN row: candle's argmin would also return 0
    if (tid == 0) s_code[q] = bi, codes[(long long)b * n_q + q0 + q] = (uint32_t)bi;
    __syncthreads();
  }
  // residual update, one subtraction per codebook in codebook order (quantization.rs:225-227)
  float* r = residual + (long long)b * res_stride;
  for (int d = tid; d < qd; d += 256) {
    float x = r[d];
    for (int q = 0; q < g; ++q) x -= emb[((long long)q * bins + s_code[q]) * qd + d];
    r[d] = x;
    if (planes) {
      const long long o = (long long)b * res_stride + d;
      uint16_t ph, pm, pl;
      split3_bf16(x, ph, pm, pl);
      reinterpret_cast<uint16_t*>(planes)[o] = ph;
      reinterpret_cast<uint16_t*>(planes)[plane_stride + o] = pm;
      reinterpret_cast<uint16_t*>(planes)[2 * plane_stride + o] = pl;
    }
  }
}

bool rvq_chain_supported(int bins, int g) { return bins % 4 == 0 && bins <= 4096 && g >= 1 && g <= DSM_RVQ_MAX_GROUP; }

void launch_rvq_chain(const float* scores, int B, int g, int bins, const float* gram, const float* emb, int qd, float* residual,
                      long long res_stride, uint32_t* codes, int n_q, int q0, __nv_bfloat16* planes, long long plane_stride,
                      cudaStream_t st) {
  if (bins <= 2048)
    launch_k(rvq_chain_kernel<DSM_RVQ_MAX_GROUP, 2>, B, 256, 0, st, scores, g, bins, gram, emb, qd, residual, res_stride, codes, n_q, q0,
             planes, plane_stride);
  else
    launch_k(rvq_chain_kernel<DSM_RVQ_MAX_GROUP, 4>, B, 256, 0, st, scores, g, bins, gram, emb, qd, residual, res_stride, codes, n_q, q0,
             planes, plane_stride);
  check_launch("rvq_chain");
}

}  // namespace dsm
