"""Oracle: batched streaming transformer.  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Restates /root/reference/server/rust/moshi/moshi-core/src/batched_transformer.rs and the rope /
layer-scale pieces of transformer.rs (:84-100, :366-403), in float32 (the reference's CPU dtype,
moshi-server/src/utils.rs:386-395).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

from .kv_cache import ScatteredCacheBuilder, ScatteredKvCache


def rope_i(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """candle_nn::rotary_emb::rope_i (interleaved pairs), called from transformer.rs:372-378.
    x ``[B, H, T, D]``; cos/sin ``[B, T, D/2]``."""
    x0, x1 = x[..., 0::2], x[..., 1::2]
    c, s = cos.unsqueeze(1), sin.unsqueeze(1)
    y0 = x0 * c - x1 * s
    y1 = x0 * s + x1 * c
    return torch.stack([y0, y1], dim=-1).flatten(-2)


def rope_tables(head_dim: int, theta: float, pos: torch.Tensor):
    """transformer.rs:385-403: inv_freq_i = 1 / theta^(i/dim) for i = 0, 2, 4, ...; freqs = pos * inv_freq
    in float32."""
    idx = torch.arange(0, head_dim, 2, dtype=torch.float32)
    inv_freq = 1.0 / torch.pow(torch.tensor(theta, dtype=torch.float32), idx / head_dim)
    freqs = pos.to(torch.float32).unsqueeze(-1) * inv_freq
    return freqs.cos(), freqs.sin()


def rms_norm(x: torch.Tensor, alpha: torch.Tensor, eps: float) -> torch.Tensor:
    """candle_nn::ops::rms_norm (batched_transformer.rs:194-198): x / sqrt(mean(x^2) + eps) * alpha."""
    ms = x.float().square().mean(dim=-1, keepdim=True)
    return x * torch.rsqrt(ms + eps) * alpha


class TransformerConfig:
    def __init__(self, d_model, num_heads, num_layers, dim_feedforward, context, max_period, norm,
                 gating: bool, layer_scale: bool, eps: float):
        self.d_model, self.num_heads, self.num_layers = d_model, num_heads, num_layers
        self.dim_feedforward, self.context, self.max_period = dim_feedforward, context, max_period
        self.norm, self.gating, self.layer_scale, self.eps = norm, gating, layer_scale, eps


class Layer:
    """batched_transformer.rs:264-368 (norm_first, no cross-attention)."""

    def __init__(self, cfg: TransformerConfig, w: Dict[str, torch.Tensor], p: str, cache: ScatteredKvCache):
        self.cfg, self.cache = cfg, cache
        self.in_proj = w[p + ".self_attn.in_proj_weight"]
        self.out_proj = w[p + ".self_attn.out_proj.weight"]
        if cfg.norm == "rms":
            self.n1 = (w[p + ".norm1.alpha"].reshape(-1), None)
            self.n2 = (w[p + ".norm2.alpha"].reshape(-1), None)
        else:
            def ln(n):
                a = w[f"{p}.{n}.alpha"].reshape(-1) if f"{p}.{n}.alpha" in w else w[f"{p}.{n}.weight"]
                return (a, w[f"{p}.{n}.bias"])
            self.n1, self.n2 = ln("norm1"), ln("norm2")
        if cfg.gating:
            self.lin_in = w[p + ".gating.linear_in.weight"]
            self.lin_out = w[p + ".gating.linear_out.weight"]
        else:
            self.lin1 = w[p + ".linear1.weight"]
            self.lin2 = w[p + ".linear2.weight"]
        self.ls1 = w[p + ".layer_scale_1.scale"] if cfg.layer_scale else None
        self.ls2 = w[p + ".layer_scale_2.scale"] if cfg.layer_scale else None

    def _norm(self, x, n):
        if self.cfg.norm == "rms":
            return rms_norm(x, n[0], self.cfg.eps)
        return F.layer_norm(x, (x.shape[-1],), n[0], n[1], self.cfg.eps)

    def _attn(self, x, cos, sin, indices, mask):
        # batched_transformer.rs:64-121
        b, t, hd = x.shape
        h = self.cfg.num_heads
        d = hd // h
        qkv = F.linear(x, self.in_proj).reshape(b, t, 3, h, d)
        q = qkv[:, :, 0].transpose(1, 2)
        k = qkv[:, :, 1].transpose(1, 2)
        v = qkv[:, :, 2].transpose(1, 2)
        q = rope_i(q, cos, sin)
        k = rope_i(k, cos, sin)
        k, v = self.cache.append(k, v, indices)
        pre = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
        pre = pre + mask
        ws = torch.softmax(pre, dim=-1)
        xs = torch.matmul(ws, v)
        xs = xs.transpose(1, 2).reshape(b, t, hd)
        return F.linear(xs, self.out_proj)

    def _mlp(self, x):
        # batched_transformer.rs:166-178
        if self.cfg.gating:
            y = F.linear(x, self.lin_in)
            b, t, _ = y.shape
            y = y.reshape(b, t, 2, -1)
            y = F.silu(y[:, :, 0]) * y[:, :, 1]
            return F.linear(y, self.lin_out)
        return F.linear(F.gelu(F.linear(x, self.lin1)), self.lin2)

    def forward(self, x, cos, sin, indices, mask):
        # batched_transformer.rs:336-363
        a = self._attn(self._norm(x, self.n1), cos, sin, indices, mask)
        if self.ls1 is not None:
            a = a * self.ls1
        x = x + a
        m = self._mlp(self._norm(x, self.n2))
        if self.ls2 is not None:
            m = m * self.ls2
        return x + m


class StreamingTransformer:
    """batched_transformer.rs:370-513."""

    def __init__(self, batch_size: int, cfg: TransformerConfig, w: Dict[str, torch.Tensor], prefix: str):
        self.cfg = cfg
        self.builder = ScatteredCacheBuilder(batch_size, cfg.context)
        d = cfg.d_model // cfg.num_heads
        self.layers = [
            Layer(cfg, w, f"{prefix}.layers.{l}", ScatteredKvCache(batch_size, cfg.num_heads, cfg.context, d))
            for l in range(cfg.num_layers)
        ]

    def forward(self, x: torch.Tensor, mask: Sequence[bool]) -> torch.Tensor:
        # batched_transformer.rs:425-459.  NOTE the rope positions are read AFTER indices_and_mask
        # advanced them (:440 then :444), i.e. an active stream's tokens are rotated by
        # (p0 + t) + s, an inactive stream's by p0 + s.  Relative phases, hence attention, are
        # unaffected for always-active streams; we reproduce the reference as written.
        b, t, _ = x.shape
        assert b == self.builder.batch_size
        indices, amask = self.builder.indices_and_mask(t, mask)
        pos = self.builder.positions.view(b, 1) + torch.arange(t).view(1, t)
        cos, sin = rope_tables(self.cfg.d_model // self.cfg.num_heads, self.cfg.max_period, pos)
        for layer in self.layers:
            x = layer.forward(x, cos, sin, indices, amask)
        return x

    def reset_state(self):
        self.builder.reset()

    def reset_batch_idx(self, b: int):
        self.builder.reset_batch_index(b)
