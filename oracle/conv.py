"""Oracle: streaming convolutions.  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Restates /root/reference/server/rust/moshi/moshi-core/src/conv.rs.  Tensors are NCL ``[B, C, T]``
float32 like the reference.  A stream tensor is ``None`` or a tensor (streaming.rs:11); a mask is
``None`` or a bool tensor ``[B]`` (streaming.rs:20,42-47).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F


def fold_weight_norm(w: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """conv.rs:27-45 / 132-141: ``weight`` if present, else ``v * g / ||v||`` with the norm over
    dims (1, 2)."""
    if prefix + ".weight" in w:
        return w[prefix + ".weight"]
    g = w[prefix + ".weight_g"]
    v = w[prefix + ".weight_v"]
    norm_v = v.square().sum(dim=(1, 2), keepdim=True).sqrt()
    return v * g / norm_v


def _pad1d(xs: torch.Tensor, pad_l: int, pad_r: int, mode: str) -> torch.Tensor:
    # conv.rs:210-216
    if mode == "constant":
        return F.pad(xs, (pad_l, pad_r))
    if mode == "replicate":
        return F.pad(xs, (pad_l, pad_r), mode="replicate")
    raise ValueError(mode)


def _masked_state(new: Optional[torch.Tensor], old: Optional[torch.Tensor], mask: Optional[torch.Tensor]):
    """The mask only gates the state update (conv.rs:347-367, 478-498)."""
    if mask is None:
        return new
    if new is None and old is None:
        return new
    if new is not None and old is None:
        return torch.where(mask.view(-1, 1, 1), new, torch.zeros_like(new))
    if new is None:
        raise RuntimeError("streaming conv should only be used with constant steps")
    if new.shape != old.shape:
        raise RuntimeError("streaming conv should only be used with constant steps")
    return torch.where(mask.view(-1, 1, 1), new, old)


class StreamableConv1d:
    """conv.rs:226-371 (causal only; groups == 1)."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int = 1, dilation: int = 1,
                 pad_mode: str = "constant"):
        self.weight, self.bias = weight, bias
        self.stride, self.dilation, self.pad_mode = stride, dilation, pad_mode
        self.kernel_size = weight.shape[-1]
        self.state_prev_xs: Optional[torch.Tensor] = None
        self.left_pad_applied = False

    @property
    def k_eff(self) -> int:
        return (self.kernel_size - 1) * self.dilation + 1

    def _conv(self, xs):
        return F.conv1d(xs, self.weight, self.bias, stride=self.stride, dilation=self.dilation)

    def forward(self, xs: torch.Tensor) -> torch.Tensor:
        # conv.rs:284-304 with get_extra_padding_for_conv1d (:197-208)
        k = self.k_eff
        padding_total = k - self.stride
        length = xs.shape[-1]
        n_frames = max(length + padding_total - k, 0) / self.stride + 1.0
        import math
        ideal = max((math.ceil(n_frames) - 1) * self.stride + k - padding_total, 0)
        extra = max(ideal - length, 0)
        return self._conv(_pad1d(xs, padding_total, extra, self.pad_mode))

    def reset_state(self):
        self.state_prev_xs = None
        self.left_pad_applied = False

    def reset_batch_idx(self, b: int):
        # conv.rs:274-281
        if self.state_prev_xs is not None:
            self.state_prev_xs = self.state_prev_xs.clone()
            self.state_prev_xs[b].zero_()

    def step(self, xs: Optional[torch.Tensor], mask: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        # conv.rs:312-370
        if xs is None:
            return None
        if not self.left_pad_applied:
            self.left_pad_applied = True
            xs = _pad1d(xs, self.k_eff - self.stride, 0, self.pad_mode)
        stride, kernel = self.stride, self.k_eff
        if self.state_prev_xs is not None:
            xs = torch.cat([self.state_prev_xs, xs], dim=-1)
        seq_len = xs.shape[-1]
        num_frames = max(seq_len + stride - kernel, 0) // stride
        if num_frames > 0:
            offset = num_frames * stride
            new_state = xs[..., offset:] if seq_len > offset else None
            in_l = (num_frames - 1) * stride + kernel
            ys = self._conv(xs[..., :in_l])
        else:
            new_state, ys = xs, None
        self.state_prev_xs = _masked_state(new_state, self.state_prev_xs, mask)
        return ys


class StreamableConvTranspose1d:
    """conv.rs:373-502 (causal; ``weight`` is ``[C_in, C_out/groups, k]``).  A depthwise weight
    (groups == C_out == C_in) is expanded to a dense one like conv.rs:144-150."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int, groups: int = 1):
        cin = weight.shape[0]
        if groups > 1:
            assert groups == cin and weight.shape[1] == 1
            eye = torch.eye(cin, dtype=weight.dtype)
            weight = weight.repeat(1, cin, 1) * eye.unsqueeze(2)
        self.weight, self.bias, self.stride = weight, bias, stride
        self.kernel_size = weight.shape[-1]
        self.state_prev_ys: Optional[torch.Tensor] = None

    def _convtr(self, xs):
        return F.conv_transpose1d(xs, self.weight, self.bias, stride=self.stride)

    def forward(self, xs):
        # conv.rs:425-441, trim_right_ratio == 1
        ys = self._convtr(xs)
        pad = max(self.kernel_size - self.stride, 0)
        return ys[..., : ys.shape[-1] - pad]

    def reset_state(self):
        self.state_prev_ys = None

    def reset_batch_idx(self, b: int):
        if self.state_prev_ys is not None:
            self.state_prev_ys = self.state_prev_ys.clone()
            self.state_prev_ys[b].zero_()

    def step(self, xs, mask):
        # conv.rs:448-501
        if xs is None:
            return None
        ys = self._convtr(xs)
        ot = ys.shape[-1]
        if self.state_prev_ys is not None:
            prev = self.state_prev_ys
            pt = prev.shape[-1]
            if self.bias is not None:
                prev = prev - self.bias.view(1, -1, 1)
            ys = torch.cat([ys[..., :pt] + prev, ys[..., pt:]], dim=-1)
        invalid = self.kernel_size - self.stride
        lhs_len = min(ot, ot - invalid)
        out = ys[..., :lhs_len] if lhs_len > 0 else None
        new_state = ys[..., lhs_len:] if ot - lhs_len > 0 else None
        self.state_prev_ys = _masked_state(new_state, self.state_prev_ys, mask)
        return out


def conv_downsample1d(weight: torch.Tensor, stride: int) -> StreamableConv1d:
    """conv.rs:505-534: learnt, k = 2*stride, no bias, replicate pad."""
    return StreamableConv1d(weight, None, stride=stride, pad_mode="replicate")


def conv_tr_upsample1d(weight: torch.Tensor, stride: int) -> StreamableConvTranspose1d:
    """conv.rs:558-585: learnt depthwise conv-transpose, k = 2*stride, no bias."""
    return StreamableConvTranspose1d(weight, None, stride=stride, groups=weight.shape[0])
