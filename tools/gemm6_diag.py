"""Where are the wrong elements of a racy tcgen05 codec GEMM run? (debug aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsm_b200 import _lib
from dsm_b200._lib import check, ptr
lib = _lib.load()
rng = np.random.default_rng(0)
B, T, C, k, s, N = [int(x) for x in sys.argv[1:7]]
S = k - s; T_in = T * s
buf = rng.standard_normal((B, S + T_in, C), dtype=np.float32)
W = (rng.standard_normal((N, k * C), dtype=np.float32) / np.sqrt(k * C)).astype(np.float32)
bias = np.zeros(N, np.float32)
ref = np.zeros((B * T, N), np.float32)
check(lib.dsm_debug_gemm_f32(0, B, T, (S + T_in) * C, s * C, N, k * C, ptr(buf), buf.size, ptr(W), ptr(bias), ptr(ref)))
flat = buf.reshape(B, -1)
K = k * C
for r in range(4):
    out = np.zeros((B * T, N), np.float32)
    check(lib.dsm_debug_gemm_f32(1, B, T, (S + T_in) * C, s * C, N, K, ptr(buf), buf.size, ptr(W), ptr(bias), ptr(out)))
    err = np.abs(out - ref)
    bad = err > 1e-4 * np.abs(ref).max()
    print(f"run {r}: {bad.sum()} bad of {bad.size}; bad rows {np.unique(np.nonzero(bad)[0])[:20]} ; bad col tiles(64) {np.unique(np.nonzero(bad)[1] // 64)[:40]}")
    if bad.any():
        i, j = np.argwhere(bad)[0]
        b_, t_ = divmod(i, T)
        a = flat[b_, t_ * s * C: t_ * s * C + K].astype(np.float64)
        contrib = (a * W[j].astype(np.float64)).reshape(-1, 64).sum(1)
        d = float(out[i, j]) - float(ref[i, j])
        close = np.argsort(np.abs(contrib + d))[:3], np.argsort(np.abs(contrib - d))[:3]
        print(f"   elem ({i},{j}) diff {d:+.5f}; k-block contribs closest to -diff {close[0]} ({contrib[close[0]]}), +diff {close[1]} ({contrib[close[1]]})")
