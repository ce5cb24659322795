#include "gemm_simt.cuh"

namespace dsm {

template <typename TA, typename TW, int BM, int BN, int TM, int TN>
static void launch_cfg(const GemmParams& p, cudaStream_t st) {
  dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM));
  launch_k(gemm_simt_kernel<TA, TW, BM, BN, TM, TN>, grid, (BM / TM) * (BN / TN), 0, st, p);
  check_launch("gemm_simt");
}

template <typename TA, typename TW>
void launch_gemm_simt(const GemmParams& p, cudaStream_t st) {
  DSM_CHECK(p.K % 16 == 0, DSM_ERR_SHAPE, "gemm_simt: K=%d is not a multiple of 16", p.K);
  DSM_CHECK(p.M > 0 && p.N > 0, DSM_ERR_SHAPE, "gemm_simt: empty problem %dx%d", p.M, p.N);
  auto ctas = [&](int bm, int bn) { return (long long)ceil_div(p.M, bm) * ceil_div(p.N, bn); };
  const long long two_waves = 2 * 148;
  if (p.N <= 32) {
    if (ctas(256, 32) >= two_waves) return launch_cfg<TA, TW, 256, 32, 8, 4>(p, st);
    return launch_cfg<TA, TW, 32, 32, 4, 4>(p, st);
  }
  if (p.N <= 64) {
    if (ctas(128, 64) >= two_waves) return launch_cfg<TA, TW, 128, 64, 8, 4>(p, st);
    if (ctas(64, 64) >= 148) return launch_cfg<TA, TW, 64, 64, 4, 4>(p, st);
    return launch_cfg<TA, TW, 32, 32, 4, 4>(p, st);
  }
  if (ctas(128, 128) >= two_waves) return launch_cfg<TA, TW, 128, 128, 8, 8>(p, st);
  if (ctas(128, 64) >= two_waves) return launch_cfg<TA, TW, 128, 64, 8, 4>(p, st);
  if (ctas(64, 64) >= 148) return launch_cfg<TA, TW, 64, 64, 4, 4>(p, st);
  return launch_cfg<TA, TW, 32, 32, 4, 4>(p, st);
}

template void launch_gemm_simt<float, float>(const GemmParams&, cudaStream_t);
template void launch_gemm_simt<__nv_bfloat16, __nv_bfloat16>(const GemmParams&, cudaStream_t);

}  // namespace dsm
