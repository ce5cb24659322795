"""Independent cross-check of the oracle's Mimi encoder architecture: the HF ``transformers`` MimiModel
(a separate implementation of the same codec) fed the SAME random weights must produce the same
pre-quantiser latents and the same RVQ codes as the oracle's streaming encode_step.

What this pins (beyond the reference's own tests): the SEANet layer order / strides / causal padding, ELU
placement, resnet skips, the codec transformer (LayerNorm, RoPE, LayerScale, GELU MLP), the replicate-padded
stride-2 downsample, the split RVQ (semantic + acoustic quantising the same latent, residual updates).
Known differences handled here: HF fuses nothing and rotates half-split RoPE pairs (we permute q/k rows),
and HF picks codes with cdist-argmin instead of argmin(c2 - x.e) (same minimiser away from ties).
"""
import numpy as np
import pytest
import torch

from helpers import MimiSpec, frame, mimi_weights, synth_audio

transformers = pytest.importorskip("transformers")


def _hf_state_dict(spec, w, decoder=False):
    t = {k: torch.from_numpy(v) for k, v in w.items()}
    sd = {}
    if decoder:
        for c in spec.decoder_layers():
            kind = "conv.conv" if c["op"] == "conv" else "convtr.convtr"
            src, dst = f"decoder.model.{c['name']}.{kind}", f"decoder.layers.{c['name']}.conv"
            sd[dst + ".weight"], sd[dst + ".bias"] = t[src + ".weight"], t[src + ".bias"]
        sd["upsample.conv.weight"] = t["upsample.convtr.convtr.convtr.weight"]
    for c in spec.encoder_convs():
        src = f"encoder.model.{c['name']}.conv.conv"
        dst = f"encoder.layers.{c['name']}.conv"
        sd[dst + ".weight"], sd[dst + ".bias"] = t[src + ".weight"], t[src + ".bias"]
    d, h = spec.dim, spec.num_heads
    hd = d // h
    # interleaved (x0,x1),(x2,x3).. pairs -> half-split layout [x0,x2,..., x1,x3,...] within each head
    perm = torch.cat([torch.arange(0, hd, 2), torch.arange(1, hd, 2)])
    rows = (torch.arange(h).view(h, 1) * hd + perm.view(1, hd)).reshape(-1)
    for name, l in [(n, l) for n in (["encoder_transformer"] + (["decoder_transformer"] if decoder else []))
                    for l in range(spec.num_layers)]:
        p, q = f"{name}.transformer.layers.{l}", f"{name}.layers.{l}"
        wq, wk, wv = t[p + ".self_attn.in_proj_weight"].split(d, dim=0)
        sd[q + ".self_attn.q_proj.weight"], sd[q + ".self_attn.k_proj.weight"] = wq[rows], wk[rows]
        sd[q + ".self_attn.v_proj.weight"] = wv
        sd[q + ".self_attn.o_proj.weight"] = t[p + ".self_attn.out_proj.weight"]
        sd[q + ".mlp.fc1.weight"], sd[q + ".mlp.fc2.weight"] = t[p + ".linear1.weight"], t[p + ".linear2.weight"]
        sd[q + ".input_layernorm.weight"], sd[q + ".input_layernorm.bias"] = t[p + ".norm1.weight"], t[p + ".norm1.bias"]
        sd[q + ".post_attention_layernorm.weight"] = t[p + ".norm2.weight"]
        sd[q + ".post_attention_layernorm.bias"] = t[p + ".norm2.bias"]
        sd[q + ".self_attn_layer_scale.scale"] = t[p + ".layer_scale_1.scale"]
        sd[q + ".mlp_layer_scale.scale"] = t[p + ".layer_scale_2.scale"]
    sd["downsample.conv.weight"] = t["downsample.conv.conv.conv.weight"]
    for ours, hf, n in (("rvq_first", "semantic_residual_vector_quantizer", 1),
                        ("rvq_rest", "acoustic_residual_vector_quantizer", spec.n_q - 1)):
        sd[f"quantizer.{hf}.input_proj.weight"] = t[f"quantizer.{ours}.input_proj.weight"]
        sd[f"quantizer.{hf}.output_proj.weight"] = t[f"quantizer.{ours}.output_proj.weight"]
        for i in range(n):
            a, b = f"quantizer.{ours}.vq.layers.{i}._codebook", f"quantizer.{hf}.layers.{i}.codebook"
            sd[b + ".initialized"] = t[a + "._initialized"]
            sd[b + ".cluster_usage"], sd[b + ".embed_sum"] = t[a + ".cluster_usage"], t[a + ".embedding_sum"]
    return sd


def test_oracle_encoder_matches_hf_mimi():
    from transformers import MimiConfig, MimiModel
    from oracle.mimi import Mimi
    torch.manual_seed(0)
    spec = MimiSpec(num_layers=2)                      # full v0_1 widths, 2 transformer layers to stay quick
    w = mimi_weights(spec, seed=0, weight_norm=False, layer_scale=0.5, decoder=False)
    cfg = MimiConfig(num_hidden_layers=spec.num_layers)
    hf = MimiModel(cfg).eval()
    missing, unexpected = hf.load_state_dict(_hf_state_dict(spec, w), strict=False)
    assert not unexpected
    assert all(k.startswith(("decoder", "upsample")) for k in missing), missing

    frames = 4
    audio = synth_audio(1, frames)
    with torch.no_grad():
        x = torch.from_numpy(audio).unsqueeze(1)
        emb = hf.encoder(x)
        emb = hf.encoder_transformer(emb.transpose(1, 2))[0].transpose(1, 2)
        hf_latent = hf.downsample(emb)                                  # [1, 512, frames]
        hf_codes = hf.quantizer.encode(hf_latent, spec.n_q).transpose(0, 1)   # [1, n_q, frames]

    o = Mimi(spec, w, 1, decoder=False)
    lat, codes = [], []
    for i in range(frames):
        c = o.encode_step(torch.from_numpy(frame(audio, i)).unsqueeze(1), None)
        lat.append(o._last_latent.copy())
        codes.append(c[:, :, 0].numpy())
    lat = np.stack(lat, axis=-1)                                        # [1, 512, frames]
    codes = np.stack(codes, axis=-1)
    err = np.abs(lat - hf_latent.numpy()).max() / np.abs(hf_latent.numpy()).max()
    assert err < 1e-4, f"latents differ from the HF implementation: rel {err:.3e}"
    agree = (codes == hf_codes.numpy()).mean()
    assert agree > 0.99, f"RVQ codes agree with the HF implementation on only {100 * agree:.1f}%"


def test_oracle_decoder_matches_hf_mimi():
    """decode_step streamed frame by frame == HF's one-shot decode of the same codes: pins the RVQ
    dequantisation, the depthwise conv-transpose upsample, the decoder transformer and the conv-transpose
    SEANet decoder (overlap-add streaming, conv.rs:448-501)."""
    from transformers import MimiConfig, MimiModel
    from oracle.mimi import Mimi
    spec = MimiSpec(num_layers=2)
    w = mimi_weights(spec, seed=0, weight_norm=False, layer_scale=0.5, decoder=True)
    hf = MimiModel(MimiConfig(num_hidden_layers=spec.num_layers)).eval()
    missing, unexpected = hf.load_state_dict(_hf_state_dict(spec, w, decoder=True), strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    frames = 4
    codes = torch.randint(0, spec.bins, (1, spec.n_q, frames), generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        hf_pcm = hf.decode(codes).audio_values                      # [1, 1, frames * 1920]
    o = Mimi(spec, w, 1, decoder=True)
    pcm = torch.cat([o.decode_step(codes[:, :, i:i + 1], None) for i in range(frames)], dim=-1)
    assert pcm.shape == hf_pcm.shape
    err = (pcm - hf_pcm).abs().max().item() / hf_pcm.abs().max().item()
    assert err < 1e-4, f"decoded audio differs from the HF implementation: rel {err:.3e}"
