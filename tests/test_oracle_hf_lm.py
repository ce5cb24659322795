"""Independent cross-check of the oracle's LM step: the HF ``transformers`` Kyutai speech-to-text decoder
(a separate implementation of the same delayed-streams LM) with the SAME random weights, run once over a
whole token sequence, must give the same hidden states / text logits as the oracle stepped one token at a
time over its ring KV cache.  (HF's non-cached forward applies plain causal attention -- its sliding
window only exists in the generation cache -- so the comparison uses a ring longer than the sequence;
the wrap-around of the ring is pinned separately by the reference's own kv_cache table,
tests/test_oracle_kv_cache.py.)

Pins: embedding sum over text + 32 audio codebooks, RMSNorm (eps 1e-8), fused in_proj split, interleaved
RoPE (HF rotates half-split pairs: q/k rows are permuted), sliding-window attention == ring of `context`
slots, SiLU-gated MLP with the gate in the first half, output norm and text head.
"""
import numpy as np
import pytest
import torch

from helpers import lm_weights, tiny_lm

transformers = pytest.importorskip("transformers")


def test_oracle_lm_matches_hf_kyutai_stt_decoder():
    from transformers import KyutaiSpeechToTextConfig
    from transformers.models.kyutai_speech_to_text.modeling_kyutai_speech_to_text import KyutaiSpeechToTextModel
    from oracle.lm import LmModel

    spec = tiny_lm(extra_heads_num=0, extra_heads_dim=0, context=32)   # d 256, 2 heads x 128, 2 layers
    w = lm_weights(spec, seed=0)
    t = {k: torch.from_numpy(v) for k, v in w.items()}
    V, ncb, av, d = spec.text_in_vocab_size, spec.audio_codebooks, spec.audio_vocab_size, spec.d_model
    cfg = KyutaiSpeechToTextConfig(
        hidden_size=d, num_hidden_layers=spec.num_layers, num_attention_heads=spec.num_heads,
        num_key_value_heads=spec.num_heads, head_dim=spec.head_dim, ffn_dim=2 * spec.hidden, vocab_size=V,
        codebook_vocab_size=av, num_codebooks=ncb, sliding_window=spec.context, max_position_embeddings=64,
        rms_norm_eps=spec.rms_eps, audio_pad_token_id=V + ncb * av,
        rope_parameters={"rope_theta": spec.max_period, "rope_type": "default"})
    hf = KyutaiSpeechToTextModel(cfg).eval()

    hd = spec.head_dim
    perm = torch.cat([torch.arange(0, hd, 2), torch.arange(1, hd, 2)])
    rows = (torch.arange(spec.num_heads).view(-1, 1) * hd + perm.view(1, hd)).reshape(-1)
    emb = torch.cat([t["text_emb.weight"]] + [t[f"emb.{i}.weight"] for i in range(ncb)] + [torch.zeros(1, d)])
    sd = {"embed_tokens.embed_tokens.weight": emb, "norm.weight": t["out_norm.alpha"].reshape(-1)}
    for l in range(spec.num_layers):
        p, q = f"transformer.layers.{l}", f"layers.{l}"
        wq, wk, wv = t[p + ".self_attn.in_proj_weight"].split(d, dim=0)
        sd[q + ".self_attn.q_proj.linear.weight"], sd[q + ".self_attn.k_proj.linear.weight"] = wq[rows], wk[rows]
        sd[q + ".self_attn.v_proj.linear.weight"] = wv
        sd[q + ".self_attn.o_proj.linear.weight"] = t[p + ".self_attn.out_proj.weight"]
        sd[q + ".mlp.fc1.weight"], sd[q + ".mlp.fc2.weight"] = t[p + ".gating.linear_in.weight"], t[p + ".gating.linear_out.weight"]
        sd[q + ".input_layernorm.weight"] = t[p + ".norm1.alpha"].reshape(-1)
        sd[q + ".post_attention_layernorm.weight"] = t[p + ".norm2.alpha"].reshape(-1)
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)

    B, T = 2, 20
    rng = np.random.default_rng(1)
    text = rng.integers(0, V, (B, T))
    audio = rng.integers(0, av, (B, T, ncb))
    ids = torch.from_numpy(np.concatenate([text[..., None], audio], axis=-1))
    with torch.no_grad():
        hf_h = hf(input_ids=ids, use_cache=False).last_hidden_state           # [B, T, d] after the output norm

    o = LmModel(spec, w, B)
    for s in range(T):
        logits, ys = o.forward_cond(torch.from_numpy(text[:, s]), torch.from_numpy(audio[:, s]), None, [True] * B)
        ref = hf_h[:, s]
        err = (ys[:, 0] - ref).abs().max().item() / ref.abs().max().item()
        assert err < 2e-4, f"step {s}: hidden state differs from the HF implementation (rel {err:.3e})"
        hf_logits = ref @ t["text_linear.weight"].T
        assert (logits[:, 0].argmax(-1) == hf_logits.argmax(-1)).all()
