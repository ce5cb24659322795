"""Streaming == one-shot for the oracle's conv1d / conv-transpose1d over the reference's own grid:
server/rust/moshi/moshi-core/src/conv.rs:613-723 (run_conv1d, run_conv_tr1d, tolerance 1e-5)."""
import pytest
import torch

from oracle.conv import StreamableConv1d, StreamableConvTranspose1d

CONV_GRID = [(1, 1, 1, 5), (2, 1, 1, 5), (2, 2, 1, 6), (3, 2, 1, 8), (3, 2, 2, 8)]   # conv.rs:702-706
CONVTR_GRID = [(1, 1, 5), (2, 1, 5), (3, 1, 5), (3, 2, 5)]                           # conv.rs:716-719


@pytest.mark.parametrize("step_size", [1, 2, 3])
@pytest.mark.parametrize("bias", [False, True])
@pytest.mark.parametrize("k,stride,dilation,length", CONV_GRID)
def test_conv1d_stream_equals_forward(step_size, bias, k, stride, dilation, length):
    g = torch.Generator().manual_seed(k * 100 + stride * 10 + dilation + step_size)
    w = torch.randn(3, 2, k, generator=g)
    b = torch.randn(3, generator=g) if bias else None
    conv = StreamableConv1d(w, b, stride=stride, dilation=dilation)
    xs = torch.randn(1, 2, step_size * length, generator=g)
    ys = conv.forward(xs)
    steps = []
    for i in range(length):
        y = conv.step(xs[:, :, step_size * i: step_size * (i + 1)], None)
        if y is not None:
            steps.append(y)
    ys_steps = torch.cat(steps, dim=-1)
    assert ys_steps.shape == ys.shape
    assert (ys - ys_steps).abs().max().item() <= 1e-5


@pytest.mark.parametrize("step_size", [1, 2, 3])
@pytest.mark.parametrize("bias", [False, True])
@pytest.mark.parametrize("k,stride,length", CONVTR_GRID)
def test_conv_tr1d_stream_equals_forward(step_size, bias, k, stride, length):
    g = torch.Generator().manual_seed(k * 100 + stride * 10 + step_size)
    w = torch.randn(2, 3, k, generator=g)
    b = torch.randn(3, generator=g) if bias else None
    conv = StreamableConvTranspose1d(w, b, stride=stride)
    xs = torch.randn(1, 2, step_size * length, generator=g)
    ys = conv.forward(xs)
    steps = []
    for i in range(length):
        y = conv.step(xs[:, :, step_size * i: step_size * (i + 1)], None)
        if y is not None:
            steps.append(y)
    ys_steps = torch.cat(steps, dim=-1)
    assert ys_steps.shape == ys.shape
    assert (ys - ys_steps).abs().max().item() <= 1e-5
