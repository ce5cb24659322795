"""Oracle: Mimi streaming codec (SEANet + codec transformer + split RVQ).  TEST INFRASTRUCTURE -- see
oracle/__init__.py.

Restates /root/reference/server/rust/moshi/moshi-core/src/{seanet.rs,quantization.rs,mimi.rs}.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from dsm_b200.spec import MimiSpec

from .conv import (StreamableConv1d, StreamableConvTranspose1d, conv_downsample1d, conv_tr_upsample1d,
                   fold_weight_norm)
from .transformer import StreamingTransformer, TransformerConfig


def _as_torch(w) -> Dict[str, torch.Tensor]:
    return {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))).float()
            for k, v in w.items()}


class ResnetBlock:
    """seanet.rs:34-151 with true_skip (no shortcut conv): y = x + conv_k1(ELU(conv_k3(ELU(x))))."""

    def __init__(self, convs: List[StreamableConv1d]):
        self.block = convs

    def forward(self, xs):
        ys = xs
        for c in self.block:
            ys = c.forward(F.elu(ys))
        return ys + xs

    def step(self, xs, mask):
        if xs is None:
            return None
        ys = xs
        for c in self.block:
            ys = c.step(F.elu(ys), mask)
        # StreamingBinOp (streaming.rs:234-263): lengths are equal in steady state, so a plain add
        assert ys is not None and ys.shape == xs.shape
        return ys + xs

    def modules(self):
        return self.block


class SeaNetEncoder:
    """seanet.rs:160-303."""

    def __init__(self, spec: MimiSpec, w: Dict[str, torch.Tensor]):
        self.layers = []       # list of ('conv'|'res', module, elu_before)
        convs = spec.encoder_convs()
        i = 0
        while i < len(convs):
            c = convs[i]
            if c["res"] == "open":
                c2 = convs[i + 1]
                blk = ResnetBlock([self._mk(c, w), self._mk(c2, w)])
                self.layers.append(("res", blk, False))
                i += 2
            else:
                self.layers.append(("conv", self._mk(c, w), c["elu"]))
                i += 1

    @staticmethod
    def _mk(c, w):
        p = f"encoder.model.{c['name']}.conv.conv"
        return StreamableConv1d(fold_weight_norm(w, p), w.get(p + ".bias"), stride=c["stride"])

    def _all_convs(self):
        for kind, m, _ in self.layers:
            if kind == "res":
                yield from m.modules()
            else:
                yield m

    def forward(self, xs):
        for kind, m, elu in self.layers:
            xs = m.forward(F.elu(xs) if elu else xs)
        return xs

    def step(self, xs, mask):
        for kind, m, elu in self.layers:
            if xs is None:
                return None
            xs = m.step(F.elu(xs) if elu else xs, mask)
        return xs

    def reset_state(self):
        for c in self._all_convs():
            c.reset_state()

    def reset_batch_idx(self, b):
        for c in self._all_convs():
            c.reset_batch_idx(b)


class SeaNetDecoder:
    """seanet.rs:312-468."""

    def __init__(self, spec: MimiSpec, w: Dict[str, torch.Tensor]):
        self.layers = []
        ls = spec.decoder_layers()
        i = 0
        while i < len(ls):
            c = ls[i]
            if c["res"] == "open":
                blk = ResnetBlock([self._mk(c, w), self._mk(ls[i + 1], w)])
                self.layers.append(("res", blk, False))
                i += 2
            else:
                self.layers.append((c["op"], self._mk(c, w), c["elu"]))
                i += 1

    @staticmethod
    def _mk(c, w):
        if c["op"] == "conv":
            p = f"decoder.model.{c['name']}.conv.conv"
            return StreamableConv1d(fold_weight_norm(w, p), w.get(p + ".bias"), stride=c["stride"])
        p = f"decoder.model.{c['name']}.convtr.convtr"
        return StreamableConvTranspose1d(fold_weight_norm(w, p), w.get(p + ".bias"), stride=c["stride"])

    def _all(self):
        for kind, m, _ in self.layers:
            if kind == "res":
                yield from m.modules()
            else:
                yield m

    def forward(self, xs):
        for kind, m, elu in self.layers:
            xs = m.forward(F.elu(xs) if elu else xs)
        return xs

    def step(self, xs, mask):
        for kind, m, elu in self.layers:
            if xs is None:
                return None
            xs = m.step(F.elu(xs) if elu else xs, mask)
        return xs

    def reset_state(self):
        for c in self._all():
            c.reset_state()

    def reset_batch_idx(self, b):
        for c in self._all():
            c.reset_batch_idx(b)


class EuclideanCodebook:
    """quantization.rs:72-153."""

    def __init__(self, w: Dict[str, torch.Tensor], p: str):
        usage = w[p + ".cluster_usage"].clamp_min(1e-5).unsqueeze(1)     # :91-92
        self.embedding = w[p + ".embedding_sum"] / usage                 # :93
        self.c2 = (self.embedding * self.embedding).sum(-1) / 2.0        # :95

    def encode(self, xs: torch.Tensor) -> torch.Tensor:
        # encode_slow, :122-131: argmin_j (c2_j - x . e_j)
        shape = xs.shape[:-1]
        xs = xs.reshape(-1, xs.shape[-1])
        dot = xs @ self.embedding.t()
        return (self.c2 - dot).argmin(dim=-1).reshape(shape)

    def decode(self, idx: torch.Tensor) -> torch.Tensor:
        return self.embedding[idx]                                       # :143-152


class ResidualVectorQuantizer:
    """quantization.rs:251-321 (force_projection: 1x1 conv in/out) over :196-249."""

    def __init__(self, w: Dict[str, torch.Tensor], p: str, n_q: int):
        self.input_proj = w[p + ".input_proj.weight"]
        self.output_proj = w[p + ".output_proj.weight"]
        self.layers = [EuclideanCodebook(w, f"{p}.vq.layers.{i}._codebook") for i in range(n_q)]

    def encode(self, xs: torch.Tensor) -> torch.Tensor:
        # xs [B, C, T] -> codes [B, n_q, T]
        residual = F.conv1d(xs, self.input_proj)
        codes = []
        for layer in self.layers:
            idx = layer.encode(residual.transpose(1, 2))          # VectorQuantization::encode, :182-185
            quantized = layer.decode(idx).transpose(1, 2)          # :187-194
            residual = residual - quantized
            codes.append(idx)
        return torch.stack(codes, 0).transpose(0, 1)

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        # codes [B, K, T]
        q = None
        for i, layer in enumerate(self.layers):
            e = layer.decode(codes[:, i]).transpose(1, 2)
            q = e if q is None else q + e
        return F.conv1d(q, self.output_proj)


class SplitResidualVectorQuantizer:
    """quantization.rs:323-391: first (semantic) and rest (acoustic) both quantise the same input."""

    def __init__(self, w, n_q: int):
        self.n_q = n_q
        self.first = ResidualVectorQuantizer(w, "quantizer.rvq_first", 1)
        self.rest = ResidualVectorQuantizer(w, "quantizer.rvq_rest", n_q - 1) if n_q > 1 else None

    def encode(self, xs):
        codes = self.first.encode(xs)
        if self.rest is not None:
            codes = torch.cat([codes, self.rest.encode(xs)], dim=1)
        return codes

    def decode(self, codes):
        q = self.first.decode(codes[:, :1])
        if self.rest is not None:
            q = q + self.rest.decode(codes[:, 1:])
        return q


class ProjectedTransformer:
    """batched_transformer.rs:515-603 with conv_layout and input_dim == d_model (no projections)."""

    def __init__(self, batch_size: int, spec: MimiSpec, w, prefix: str):
        cfg = TransformerConfig(spec.dim, spec.num_heads, spec.num_layers, spec.dim_feedforward, spec.context,
                                spec.max_period, norm="layer", gating=False, layer_scale=True, eps=spec.ln_eps)
        self.transformer = StreamingTransformer(batch_size, cfg, w, prefix + ".transformer")

    def step(self, xs, mask: Sequence[bool]):
        if xs is None:
            return None
        return self.transformer.forward(xs.transpose(1, 2), mask).transpose(1, 2)


class Mimi:
    """mimi.rs:95-245, batched construction (``Mimi::batched``)."""

    def __init__(self, spec: MimiSpec, weights, batch_size: int, decoder: bool = True):
        w = _as_torch(weights)
        self.spec, self.batch_size = spec, batch_size
        self.encoder = SeaNetEncoder(spec, w)
        self.encoder_transformer = ProjectedTransformer(batch_size, spec, w, "encoder_transformer")
        self.downsample = conv_downsample1d(w["downsample.conv.conv.conv.weight"], spec.downsample_stride)
        self.quantizer = SplitResidualVectorQuantizer(w, spec.n_q)
        self.has_decoder = decoder
        if decoder:
            self.decoder = SeaNetDecoder(spec, w)
            self.decoder_transformer = ProjectedTransformer(batch_size, spec, w, "decoder_transformer")
            self.upsample = conv_tr_upsample1d(w["upsample.convtr.convtr.convtr.weight"], spec.downsample_stride)

    @staticmethod
    def _mask(mask, b):
        if mask is None:
            return [True] * b, None
        m = [bool(x) for x in mask]
        return m, torch.tensor(m)

    def encode_latent_step(self, pcm: torch.Tensor, mask=None) -> Optional[torch.Tensor]:
        """encode_step up to (excluding) the quantiser: ``[B, 1, 1920]`` -> ``[B, 512, 1]``."""
        cpu_mask, tmask = self._mask(mask, pcm.shape[0])
        xs = self.encoder.step(pcm, tmask)
        self._last_seanet = None if xs is None else xs.transpose(1, 2).numpy().copy()
        xs = self.encoder_transformer.step(xs, cpu_mask)
        self._last_transformer = None if xs is None else xs.transpose(1, 2).numpy().copy()
        return self.downsample.step(xs, tmask)

    def encode_step(self, pcm: torch.Tensor, mask=None) -> Optional[torch.Tensor]:
        # mimi.rs:195-206
        xs = self.encode_latent_step(pcm, mask)
        if xs is None:
            return None
        self._last_latent = xs[:, :, 0].numpy().copy()
        return self.quantizer.encode(xs)

    def decode_step(self, codes: torch.Tensor, mask=None) -> Optional[torch.Tensor]:
        # mimi.rs:217-225
        cpu_mask, tmask = self._mask(mask, codes.shape[0])
        emb = self.quantizer.decode(codes)
        emb = self.upsample.step(emb, tmask)
        out = self.decoder_transformer.step(emb, cpu_mask)
        return self.decoder.step(out, tmask)

    def reset_state(self):
        # mimi.rs:227-234
        self.encoder.reset_state()
        self.encoder_transformer.transformer.reset_state()
        self.downsample.reset_state()
        if self.has_decoder:
            self.decoder.reset_state()
            self.decoder_transformer.transformer.reset_state()
            self.upsample.reset_state()

    def reset_batch_idx(self, b: int):
        # mimi.rs:236-244.  The reference resets encoder_transformer twice and never
        # decoder_transformer; we reset both.
        self.encoder_transformer.transformer.reset_batch_idx(b)
        self.encoder.reset_batch_idx(b)
        self.downsample.reset_batch_idx(b)
        if self.has_decoder:
            self.decoder_transformer.transformer.reset_batch_idx(b)
            self.decoder.reset_batch_idx(b)
            self.upsample.reset_batch_idx(b)
