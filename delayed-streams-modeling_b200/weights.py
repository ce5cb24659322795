"""Seeded random-init weights under the reference's tensor names and shapes.

No checkpoints are reachable offline, so every test and benchmark runs on weights generated here.
Names/shapes follow the reference's loaders so a real safetensors checkpoint drops in unchanged:
  * LM:    server/rust/moshi/moshi-core/src/lm.rs:840-883,
           batched_transformer.rs:41-45,152-160,189,289-305
  * Mimi:  mimi.rs:119-159, seanet.rs:177-245,330-399, conv.rs:35-43,131-141,
           quantization.rs:88-90,210-213,275-302,343-360

Tensors are returned as a ``dict[str, numpy.ndarray]`` (float32).  Each tensor has its own RNG stream
derived from (seed, name) so generation order does not matter.
"""
from __future__ import annotations

import zlib
from typing import Dict

import numpy as np

from .spec import LmSpec, MimiSpec


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def _normal(seed, name, shape, std):
    return (_rng(seed, name).standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)


def mimi_weights(spec: MimiSpec, seed: int = 0, weight_norm: bool = True, layer_scale: float | None = None,
                 decoder: bool = True) -> Dict[str, np.ndarray]:
    """Random Mimi weights.  With ``weight_norm`` the SEANet convs are emitted as
    ``weight_g``/``weight_v`` pairs (folded at load, conv.rs:27-45), otherwise as plain ``weight``."""
    w: Dict[str, np.ndarray] = {}
    ls = spec.layer_scale_init if layer_scale is None else layer_scale

    def conv(prefix, cout, cin, k, bias=True, transposed=False, gain=1.0):
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        fan_in = cin * k
        v = _normal(seed, prefix + ".v", shape, gain / np.sqrt(fan_in))
        if weight_norm:
            w[prefix + ".weight_v"] = v
            n0 = shape[0]
            # g is perturbed around |v| so the fold is not the identity
            norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2))).astype(np.float32)
            g = norm * (1.0 + 0.1 * _rng(seed, prefix + ".g").standard_normal(n0, dtype=np.float32))
            w[prefix + ".weight_g"] = g.reshape(n0, 1, 1).astype(np.float32)
        else:
            w[prefix + ".weight"] = v
        if bias:
            w[prefix + ".bias"] = _normal(seed, prefix + ".bias", (cout,), 0.05)

    # ELU shrinks activations; a gain > 1 keeps the latent scale O(1) through the stack
    for c in spec.encoder_convs():
        conv(f"encoder.model.{c['name']}.conv.conv", c["cout"], c["cin"], c["k"], gain=1.4)
    if decoder:
        for c in spec.decoder_layers():
            if c["op"] == "conv":
                conv(f"decoder.model.{c['name']}.conv.conv", c["cout"], c["cin"], c["k"], gain=1.4)
            else:
                conv(f"decoder.model.{c['name']}.convtr.convtr", c["cout"], c["cin"], c["k"],
                     transposed=True, gain=1.4 * np.sqrt(c["stride"]))

    d, ff = spec.dim, spec.dim_feedforward
    for tname in (["encoder_transformer"] + (["decoder_transformer"] if decoder else [])):
        for l in range(spec.num_layers):
            p = f"{tname}.transformer.layers.{l}"
            w[p + ".self_attn.in_proj_weight"] = _normal(seed, p + ".in_proj", (3 * d, d), 1 / np.sqrt(d))
            w[p + ".self_attn.out_proj.weight"] = _normal(seed, p + ".out_proj", (d, d), 1 / np.sqrt(d))
            for n in ("norm1", "norm2"):
                w[f"{p}.{n}.weight"] = (1.0 + 0.1 * _rng(seed, f"{p}.{n}.w").standard_normal(d, dtype=np.float32))
                w[f"{p}.{n}.bias"] = _normal(seed, f"{p}.{n}.b", (d,), 0.05)
            w[p + ".linear1.weight"] = _normal(seed, p + ".linear1", (ff, d), 1 / np.sqrt(d))
            w[p + ".linear2.weight"] = _normal(seed, p + ".linear2", (d, ff), 1 / np.sqrt(ff))
            w[p + ".layer_scale_1.scale"] = np.full((d,), ls, np.float32)
            w[p + ".layer_scale_2.scale"] = np.full((d,), ls, np.float32)

    w["downsample.conv.conv.conv.weight"] = _normal(
        seed, "downsample", (d, d, 2 * spec.downsample_stride), 1 / np.sqrt(d * 2 * spec.downsample_stride))
    if decoder:
        w["upsample.convtr.convtr.convtr.weight"] = _normal(
            seed, "upsample", (d, 1, 2 * spec.downsample_stride), 0.7)

    q = spec.qdim
    for part, n_layers in (("rvq_first", 1), ("rvq_rest", spec.n_q - 1)):
        p = f"quantizer.{part}"
        w[p + ".input_proj.weight"] = _normal(seed, p + ".in", (q, d, 1), 1 / np.sqrt(d))
        w[p + ".output_proj.weight"] = _normal(seed, p + ".out", (d, q, 1), 1 / np.sqrt(q))
        for l in range(n_layers):
            c = f"{p}.vq.layers.{l}._codebook"
            # codebook scale decays with depth like a trained residual quantiser's would
            sigma = 0.6 * (0.93 ** l)
            usage = 1.0 + _rng(seed, c + ".usage").random(spec.bins, dtype=np.float32) * 3.0
            emb = _normal(seed, c + ".emb", (spec.bins, q), sigma)
            w[c + "._initialized"] = np.ones((1,), np.float32)
            w[c + ".cluster_usage"] = usage.astype(np.float32)
            w[c + ".embedding_sum"] = (emb * usage[:, None]).astype(np.float32)
    return w


def lm_weights(spec: LmSpec, seed: int = 0) -> Dict[str, np.ndarray]:
    """Random LM weights (float32; the device path rounds them to bf16, the oracle is handed the
    same bf16-rounded values so weight rounding is not part of the comparison)."""
    w: Dict[str, np.ndarray] = {}
    d, h = spec.d_model, spec.hidden
    emb_std = 1.0 / np.sqrt(1 + spec.audio_codebooks)
    w["text_emb.weight"] = _normal(seed, "text_emb", (spec.text_in_vocab_size, d), emb_std)
    for i in range(spec.audio_codebooks):
        w[f"emb.{i}.weight"] = _normal(seed, f"emb.{i}", (spec.audio_vocab_size, d), emb_std)
    for l in range(spec.num_layers):
        p = f"transformer.layers.{l}"
        w[p + ".self_attn.in_proj_weight"] = _normal(seed, p + ".in_proj", (3 * d, d), 1 / np.sqrt(d))
        w[p + ".self_attn.out_proj.weight"] = _normal(seed, p + ".out_proj", (d, d), 0.5 / np.sqrt(d))
        w[p + ".norm1.alpha"] = (1.0 + 0.1 * _rng(seed, p + ".n1").standard_normal((1, 1, d), dtype=np.float32))
        w[p + ".norm2.alpha"] = (1.0 + 0.1 * _rng(seed, p + ".n2").standard_normal((1, 1, d), dtype=np.float32))
        w[p + ".gating.linear_in.weight"] = _normal(seed, p + ".lin_in", (2 * h, d), 1 / np.sqrt(d))
        w[p + ".gating.linear_out.weight"] = _normal(seed, p + ".lin_out", (d, h), 0.5 / np.sqrt(h))
    w["out_norm.alpha"] = (1.0 + 0.1 * _rng(seed, "out_norm").standard_normal((1, 1, d), dtype=np.float32))
    # a wide text head spreads the logits so greedy argmax is well conditioned under random init
    w["text_linear.weight"] = _normal(seed, "text_linear", (spec.text_out_vocab_size, d), 2.0 / np.sqrt(d))
    for i in range(spec.extra_heads_num):
        w[f"extra_heads.{i}.weight"] = _normal(seed, f"extra_heads.{i}", (spec.extra_heads_dim, d), 1 / np.sqrt(d))
    return w


def round_bf16(a: np.ndarray) -> np.ndarray:
    """float32 -> nearest-even bf16 -> float32 (what the device sees after its cast)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


def to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """float32 -> uint16 array holding nearest-even bf16 bit patterns."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) >> 16).astype(np.uint16)
