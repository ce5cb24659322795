"""GPU check of the tcgen05 GEMM (impl 1) and the CUDA-core GEMM (impl 0) against numpy float64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsm_b200 import _lib
from dsm_b200._lib import check, ptr
from dsm_b200.weights import to_bf16_bits, round_bf16

lib = _lib.load()
rng = np.random.default_rng(0)
ok = True
cases = [(256, 6144, 2048, 0), (256, 2048, 2048, 1), (256, 11264, 2048, 2), (256, 2048, 5632, 1), (256, 8000, 2048, 0),
         (3, 768, 256, 0), (3, 256, 256, 1), (3, 1408, 256, 2), (3, 256, 704, 1), (100, 256, 256, 0), (700, 384, 128, 1),
         (64, 128, 64, 0), (130, 512, 192, 2)]
for (M, N, K, mode) in cases:
    X = rng.standard_normal((M, K), dtype=np.float32)
    W = rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)
    Xr, Wr = round_bf16(X).astype(np.float64), round_bf16(W).astype(np.float64)
    ref = Xr @ Wr.T
    base = rng.standard_normal((M, N), dtype=np.float32)
    Xb, Wb = to_bf16_bits(X), to_bf16_bits(W)      # keep alive across the C call
    for impl in (0, 1):
        out = base.copy() if mode == 1 else np.zeros((M, N), np.float32)
        gate = np.zeros((M, N // 2), np.uint16)
        check(lib.dsm_debug_gemm_bf16(impl, mode, M, N, K, ptr(Xb), ptr(Wb), ptr(out), ptr(gate)))
        if mode == 2:
            h = N // 2
            r = ref[:, :h] / (1 + np.exp(-ref[:, :h])) * ref[:, h:]
            got = (gate.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
            tol = 1e-2
        else:
            r = ref + (base if mode == 1 else 0)
            got = out.astype(np.float64)
            tol = 2e-5
        err = np.abs(got - r).max() / (np.abs(r).max() + 1e-30)
        flag = "ok " if err < tol else "BAD"
        ok &= err < tol
        print(f"{flag} impl={impl} mode={mode} M={M} N={N} K={K} rel_err={err:.3e}", flush=True)
print("ALL OK" if ok else "FAILURES")
