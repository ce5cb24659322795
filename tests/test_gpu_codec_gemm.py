"""GPU checks of the tcgen05 codec GEMM (gemm_tc6) through the C ABI's debug entry point.

* accuracy against numpy float64 for dense rows and overlapping conv windows, for the f32-A path (converter
  warps) and the pre-split bf16-plane path, including the plane outputs (which must sum back bit-exactly);
* determinism: every shape is launched repeatedly on the same inputs and must return identical bits.  This
  is the regression test for the timing-dependent corruption that unpaced MMA issue produced.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# (B, T_out, C_in, k, stride, C_out): conv layers; k = 1, T_out = 1..2 -> dense linear with K = C_in
SHAPES = [(2, 2, 1024, 2, 1, 4096), (2, 96, 256, 2, 1, 640), (2, 1920, 64, 3, 1, 32), (2, 1920, 32, 1, 1, 64),
          (3, 2, 512, 7, 1, 1024), (2, 2, 2048, 1, 1, 512), (2, 96, 128, 10, 5, 256), (130, 2, 512, 16, 8, 1024),
          (64, 1920, 64, 3, 1, 32), (64, 480, 64, 8, 4, 128), (256, 1, 256, 1, 1, 2048)]


def _case(rng, B, T, C, k, s, N):
    S, T_in = k - s, T * s
    buf = rng.standard_normal((B, S + T_in, C), dtype=np.float32)
    W = (rng.standard_normal((N, k * C), dtype=np.float32) / np.sqrt(k * C)).astype(np.float32)
    bias = rng.standard_normal(N, dtype=np.float32)
    return buf, W, bias, (S + T_in) * C, s * C


def _run(lib, impl, B, T, N, K, buf, W, bias, stride_b, stride_t):
    from dsm_b200._lib import check, ptr
    out = np.zeros((B * T, N), np.float32)
    check(lib.dsm_debug_gemm_f32(impl, B, T, stride_b, stride_t, N, K, ptr(buf), buf.size, ptr(W), ptr(bias), ptr(out)))
    return out


@pytest.mark.parametrize("shape", SHAPES[:8], ids=lambda s: "x".join(map(str, s)))
def test_accuracy_against_float64(shape):
    from dsm_b200 import _lib
    lib = _lib.load()
    B, T, C, k, s, N = shape
    buf, W, bias, sb, st = _case(np.random.default_rng(1), *shape)
    flat = buf.reshape(B, -1).astype(np.float64)
    ref = np.empty((B, T, N), np.float64)
    for t in range(T):
        ref[:, t] = flat[:, t * st: t * st + k * C] @ W.astype(np.float64).T + bias
    for impl in (1, 2, 3):      # f32 A; pre-split A; pre-split A + plane outputs (checked inside the library)
        out = _run(lib, impl, B, T, N, k * C, buf, W, bias, sb, st)
        err = np.abs(out.reshape(B, T, N) - ref).max() / np.abs(ref).max()
        assert err < 5e-6, f"impl {impl}: rel err {err:.3e}"     # f32-level: the CUDA-core GEMM gives 2.7e-6 at K = 8192


@pytest.mark.parametrize("impl", [1, 2], ids=["f32_A", "presplit_A"])
def test_repeated_launches_are_bit_identical(impl):
    from dsm_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for shape in SHAPES:
        B, T, C, k, s, N = shape
        buf, W, bias, sb, st = _case(rng, *shape)
        first = _run(lib, impl, B, T, N, k * C, buf, W, bias, sb, st)
        ref = _run(lib, 0, B, T, N, k * C, buf, W, bias, sb, st)            # CUDA-core GEMM
        assert np.abs(first - ref).max() / np.abs(ref).max() < 1e-5, shape
        for _ in range(5):
            again = _run(lib, impl, B, T, N, k * C, buf, W, bias, sb, st)
            assert np.array_equal(first, again), f"{shape}: results differ between launches"


# ---- the LM's bf16 tcgen05 GEMM (gemm_tc) ---------------------------------------------------------
LM_SHAPES = [(256, 6144, 2048, 0), (256, 8000, 2048, 0), (256, 11264, 2048, 2), (3, 768, 256, 0), (64, 128, 64, 0),
             (100, 256, 256, 0), (130, 512, 192, 2), (1, 6144, 2048, 0)]


def test_lm_gemm_accuracy_and_repeated_launches():
    """Values against float64 on the bf16-rounded operands (f32 accumulation: rel 2e-5; bf16 gate output: 1e-2);
    repeated launches agree to rounding (split-K partial sums meet through f32 atomics, so the order of the
    additions, not more, may differ) -- the corruption this guards against was 1e-2 of the value range."""
    from dsm_b200 import _lib
    from dsm_b200._lib import check, ptr
    from dsm_b200.weights import round_bf16, to_bf16_bits
    lib = _lib.load()
    rng = np.random.default_rng(3)
    for (M, N, K, mode) in LM_SHAPES:
        X = rng.standard_normal((M, K), dtype=np.float32)
        W = rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)
        ref = round_bf16(X).astype(np.float64) @ round_bf16(W).astype(np.float64).T
        Xb, Wb = to_bf16_bits(X), to_bf16_bits(W)          # keep alive across the C calls
        first = None
        for _ in range(4):
            out = np.zeros((M, N), np.float32)
            gate = np.zeros((M, N // 2), np.uint16)
            check(lib.dsm_debug_gemm_bf16(1, mode, M, N, K, ptr(Xb), ptr(Wb), ptr(out), ptr(gate)))
            got = gate.copy() if mode == 2 else out
            if first is None:
                first = got
                if mode == 2:
                    h = N // 2
                    r = ref[:, :h] / (1 + np.exp(-ref[:, :h])) * ref[:, h:]
                    val = (gate.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
                    assert np.abs(val - r).max() / np.abs(r).max() < 1e-2, (M, N, K)
                else:
                    assert np.abs(out - ref).max() / np.abs(ref).max() < 2e-5, (M, N, K)
            else:
                if mode == 2:
                    a = (first.astype(np.uint32) << 16).view(np.float32)
                    b = (got.astype(np.uint32) << 16).view(np.float32)
                    tol = 1e-2          # one bf16 ulp of the output
                else:
                    a, b, tol = first, got, 1e-6
                assert np.abs(a - b).max() <= tol * np.abs(a).max(), f"{(M, N, K, mode)}: results differ between launches"
