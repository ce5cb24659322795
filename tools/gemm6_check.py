"""GPU check of the codec GEMM kernels (CUDA-core impl 0, tcgen05 split-bf16 impl 1) against numpy f64,
dense rows and overlapping conv windows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsm_b200 import _lib
from dsm_b200._lib import check, ptr

lib = _lib.load()
rng = np.random.default_rng(0)
ok = True
# (B, T_out, C_in, k, stride, C_out): conv layers; k=1,stride=1,T_out=1 -> dense linear with K = C_in
cases = [(4, 1, 512, 1, 1, 512), (6, 2, 512, 1, 1, 1536), (3, 480, 64, 8, 4, 128), (2, 1920, 64, 3, 1, 32),
         (2, 1920, 32, 1, 1, 64), (5, 96, 128, 10, 5, 256), (9, 16, 256, 12, 6, 512), (130, 2, 512, 16, 8, 1024),
         (7, 1, 256, 1, 1, 2048), (3, 16, 512, 3, 1, 256), (300, 1, 512, 4, 2, 512)]
for (B, T, C, k, s, N) in cases:
    S = k - s
    T_in = T * s
    buf = rng.standard_normal((B, S + T_in, C), dtype=np.float32)
    W = (rng.standard_normal((N, k * C), dtype=np.float32) / np.sqrt(k * C)).astype(np.float32)
    bias = rng.standard_normal(N, dtype=np.float32)
    # reference: window rows
    ref = np.empty((B, T, N), np.float64)
    flat = buf.reshape(B, -1).astype(np.float64)
    for t in range(T):
        ref[:, t] = flat[:, t * s * C: t * s * C + k * C] @ W.astype(np.float64).T + bias
    for impl in (0, 1, 2, 3):
        out = np.zeros((B * T, N), np.float32)
        check(lib.dsm_debug_gemm_f32(impl, B, T, (S + T_in) * C, s * C, N, k * C, ptr(buf), buf.size, ptr(W), ptr(bias), ptr(out)))
        err = np.abs(out.reshape(B, T, N) - ref).max() / np.abs(ref).max()
        flag = "ok " if err < 3e-6 else "BAD"
        ok &= err < 3e-6
        print(f"{flag} impl={impl} B={B} T={T} C_in={C} k={k} s={s} N={N} K={k*C} rel_err={err:.3e}", flush=True)
print("ALL OK" if ok else "FAILURES")
