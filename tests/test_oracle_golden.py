"""The oracle reproduces the committed golden vectors (tests/golden/make_golden.py), and satisfies the
size-independent properties the reference's own conv tests check (streaming == one-shot,
conv.rs:613-723) on the full SEANet encoder stack."""
import numpy as np
import torch

from helpers import GOLDEN, frame, golden_masks, lm_weights, mimi_weights, oracle_lm_weights, synth_audio, tiny_lm, tiny_mimi
from oracle.lm import AsrState, LmModel
from oracle.mimi import Mimi, SeaNetEncoder, _as_torch


def test_mimi_oracle_matches_golden():
    torch.set_num_threads(1)
    g = np.load(f"{GOLDEN}/mimi_tiny.npz")
    spec = tiny_mimi()
    w = mimi_weights(spec, seed=0, layer_scale=0.5)
    enc, dec = Mimi(spec, w, 2, decoder=False), Mimi(spec, w, 2, decoder=True)
    audio = synth_audio(2, len(golden_masks()))
    for i, m in enumerate(golden_masks()):
        c = enc.encode_step(torch.from_numpy(frame(audio, i)).unsqueeze(1), m)
        assert (c[:, :, 0].numpy() == g["codes"][i]).all()
        np.testing.assert_allclose(enc._last_latent, g["latents"][i], rtol=0, atol=2e-4)
        p = dec.decode_step(c, m)
        np.testing.assert_allclose(p[:, 0].numpy(), g["pcm"][i], rtol=0, atol=2e-3)


def test_asr_oracle_matches_golden():
    torch.set_num_threads(1)
    g = np.load(f"{GOLDEN}/asr_tiny.npz")
    mspec, lspec = tiny_mimi(), tiny_lm(text_in_vocab_size=9, text_out_vocab_size=8)
    asr = AsrState(2, lspec.asr_delay_in_tokens, Mimi(mspec, mimi_weights(mspec, seed=0, layer_scale=0.5), 2, decoder=False),
                   LmModel(lspec, oracle_lm_weights(lm_weights(lspec, seed=0)), 2))
    audio = synth_audio(2, len(golden_masks()))
    for i, m in enumerate(golden_masks()):
        out = asr.step_pcm(torch.from_numpy(frame(audio, i)).unsqueeze(1), m)
        assert (asr.last_tokens.numpy() == g["tokens"][i]).all()
        assert repr([(x[0], x[1]) if x[0] == "step" else x for x in out]) == str(g["msgs"][i])


def test_seanet_encoder_stream_equals_forward():
    """Streaming in 1920-sample steps == one-shot forward on the concatenated audio (causal stack)."""
    spec = tiny_mimi()
    w = _as_torch(mimi_weights(spec, seed=1))
    enc = SeaNetEncoder(spec, w)
    audio = torch.from_numpy(synth_audio(1, 3)).unsqueeze(1)
    ref = enc.forward(audio)
    enc.reset_state()
    outs = [enc.step(audio[:, :, i * 1920:(i + 1) * 1920], None) for i in range(3)]
    got = torch.cat(outs, dim=-1)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()


def test_rvq_encode_decode_reduces_residual():
    """decode(encode(x)) approaches x as codebooks are added (residual quantisation property)."""
    spec = tiny_mimi()
    m = Mimi(spec, mimi_weights(spec, seed=0), 1, decoder=False)
    x = torch.randn(1, 256, 1, generator=torch.Generator().manual_seed(0)) * 2.0
    residual = x.clone()
    norms = [residual.norm().item()]
    for layer in m.quantizer.rest.layers:
        idx = layer.encode(residual.transpose(1, 2))
        residual = residual - layer.decode(idx).transpose(1, 2)
        norms.append(residual.norm().item())
    assert all(b <= a + 1e-6 for a, b in zip(norms, norms[1:]))
