"""Race hunt for the tcgen05 codec GEMM: every case is launched many times on the same inputs; any run that
differs from the first one (bitwise) is reported.  Shapes follow the Mimi encoder and decoder layers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsm_b200 import _lib
from dsm_b200._lib import check, ptr

lib = _lib.load()
rng = np.random.default_rng(0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 1     # 1: f32 A through the converter warps, 2: pre-split A planes
# (B, T_out, C_in, k, stride, C_out)
cases = [(2, 2, 1024, 2, 1, 4096), (2, 16, 512, 2, 1, 1536), (2, 96, 256, 2, 1, 640), (2, 480, 128, 2, 1, 256),
         (2, 1920, 64, 3, 1, 32), (2, 1920, 32, 1, 1, 64), (3, 2, 512, 7, 1, 1024), (2, 2, 512, 1, 1, 2048),
         (2, 2, 2048, 1, 1, 512), (2, 480, 64, 8, 4, 128), (2, 96, 128, 10, 5, 256), (2, 16, 256, 12, 6, 512),
         (2, 2, 512, 16, 8, 1024), (64, 1920, 64, 3, 1, 32), (64, 480, 64, 8, 4, 128)]
bad = 0
for (B, T, C, k, s, N) in cases:
    S = k - s
    T_in = T * s
    buf = rng.standard_normal((B, S + T_in, C), dtype=np.float32)
    W = (rng.standard_normal((N, k * C), dtype=np.float32) / np.sqrt(k * C)).astype(np.float32)
    bias = rng.standard_normal(N, dtype=np.float32)
    ref = np.zeros((B * T, N), np.float32)
    check(lib.dsm_debug_gemm_f32(0, B, T, (S + T_in) * C, s * C, N, k * C, ptr(buf), buf.size, ptr(W), ptr(bias), ptr(ref)))
    first = None
    ndiff = 0
    worst = 0.0
    for r in range(reps):
        out = np.zeros((B * T, N), np.float32)
        check(lib.dsm_debug_gemm_f32(impl, B, T, (S + T_in) * C, s * C, N, k * C, ptr(buf), buf.size, ptr(W), ptr(bias), ptr(out)))
        if first is None:
            first = out
        elif not np.array_equal(first, out):
            ndiff += 1
        worst = max(worst, float(np.abs(out - ref).max() / np.abs(ref).max()))
    tag = "ok " if ndiff == 0 and worst < 1e-5 else "BAD"
    bad += tag == "BAD"
    print(f"{tag} B={B} T={T} C_in={C} k={k} s={s} N={N} K={k*C}: {ndiff}/{reps-1} runs differ, worst rel err vs CUDA-core {worst:.3e}", flush=True)
print("ALL OK" if not bad else "FAILURES")
