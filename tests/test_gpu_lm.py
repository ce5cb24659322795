"""GPU parity of the LM step (LmModel::forward_cond) and of its tcgen05 projection GEMMs.

Bar: logits within bf16 tolerance (rel 1e-2) of the f32 oracle fed the same bf16-rounded weights; greedy
argmax identical."""
import numpy as np
import pytest
import torch

from helpers import LmSpec, active, lm_weights, oracle_lm_weights, random_masks, rel_err, tiny_lm

pytestmark = pytest.mark.gpu


def _run(spec, b, steps, seed=3):
    from dsm_b200 import api
    from oracle.lm import LmModel as OLm
    w = lm_weights(spec, seed=0)
    o = OLm(spec, oracle_lm_weights(w), b)
    g = api.LmModel(spec, w, b)
    rng = np.random.default_rng(seed)
    for i, m in enumerate(random_masks(steps, b, seed=seed, first_active=3)):
        text = rng.integers(0, spec.text_in_vocab_size, b)
        audio = rng.integers(0, spec.audio_vocab_size, (b, spec.audio_codebooks))
        cond = rng.standard_normal(spec.d_model).astype(np.float32) * 0.1 if i % 4 == 3 else None
        ol, oy = o.forward_cond(torch.from_numpy(text), torch.from_numpy(audio),
                                None if cond is None else torch.from_numpy(cond),
                                [True] * b if m is None else list(m))
        gl, gy = g.forward_cond(text, audio, cond, m)
        a = active(m, b)
        assert rel_err(gy[a], oy[:, 0].numpy()[a]) < 1e-2
        assert rel_err(gl[a], ol[:, 0].numpy()[a]) < 1e-2
        assert (gl.argmax(-1)[a] == ol[:, 0].numpy().argmax(-1)[a]).all()


def test_tiny_lm_ring_wraps():
    _run(tiny_lm(), b=3, steps=30)          # context 12 -> the ring wraps twice


def test_stt_1b_shaped_layers():
    """Two layers at the real stt-1b widths (d 2048, 16 heads x 128, gated hidden 5632, vocab 8000)."""
    spec = LmSpec(name="stt-1b-2layers", num_layers=2, context=20)
    _run(spec, b=5, steps=6)


def test_stt_2_6b_shaped_layers():
    """Two layers at the stt-2.6b widths (32 heads x 64, vocab 4000, no extra heads)."""
    spec = LmSpec(name="stt-2.6b-2layers", num_heads=32, num_layers=2, context=16, text_in_vocab_size=4001,
                  text_out_vocab_size=4000, extra_heads_num=0, extra_heads_dim=0)
    _run(spec, b=4, steps=5)


@pytest.mark.parametrize("M,N,K,mode", [(256, 6144, 2048, 0), (256, 2048, 2048, 1), (256, 11264, 2048, 2),
                                        (256, 2048, 5632, 1), (256, 8000, 2048, 0), (3, 768, 256, 0),
                                        (130, 512, 192, 2), (700, 384, 128, 1), (2000, 256, 128, 0),
                                        (2000, 11264, 256, 2), (1000, 6144, 512, 0), (1000, 2048, 512, 1),
                                        (2000, 6144, 256, 0), (2000, 2048, 512, 1), (2000, 8000, 256, 0)])   # wide tiles
def test_tcgen05_gemm_against_numpy(M, N, K, mode):
    from dsm_b200 import _lib
    from dsm_b200._lib import check, ptr
    from dsm_b200.weights import round_bf16, to_bf16_bits
    lib = _lib.load()
    rng = np.random.default_rng(M + N + K)
    X = rng.standard_normal((M, K), dtype=np.float32)
    W = rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)
    ref = round_bf16(X).astype(np.float64) @ round_bf16(W).astype(np.float64).T
    base = rng.standard_normal((M, N), dtype=np.float32)
    xb, wb = to_bf16_bits(X), to_bf16_bits(W)
    out = base.copy() if mode == 1 else np.zeros((M, N), np.float32)
    gate = np.zeros((M, max(N // 2, 1)), np.uint16)
    check(lib.dsm_debug_gemm_bf16(1, mode, M, N, K, ptr(xb), ptr(wb), ptr(out), ptr(gate)))
    if mode == 2:
        h = N // 2
        want = ref[:, :h] / (1 + np.exp(-ref[:, :h])) * ref[:, h:]
        got = (gate.astype(np.uint32) << 16).view(np.float32)
        assert rel_err(got, want) < 1e-2          # bf16 output
    else:
        want = ref + (base if mode == 1 else 0)
        assert rel_err(out, want) < 2e-5          # f32 accumulate of exact bf16 products
