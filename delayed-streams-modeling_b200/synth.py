"""Synthetic 24 kHz audio for parity tests and benchmarks (SURVEY.md §8(d) generator).

Per stream ``s``: ``0.3 * sum_{k<4} a_k sin(2 pi f_k t + phi_k) + 0.05 N(0,1)``, amplitude-modulated by
a 2-5 Hz raised cosine, clipped to [-1, 1]; ``numpy.random.default_rng(1234 + s)``.
"""
from __future__ import annotations

import numpy as np

from .spec import FRAME_SIZE, SAMPLE_RATE


def synth_audio(n_streams: int, n_frames: int, first_stream: int = 0) -> np.ndarray:
    """Returns float32 ``[n_streams, n_frames * 1920]``."""
    n = n_frames * FRAME_SIZE
    t = np.arange(n, dtype=np.float64) / SAMPLE_RATE
    out = np.empty((n_streams, n), np.float32)
    for i in range(n_streams):
        rng = np.random.default_rng(1234 + first_stream + i)
        f = rng.uniform(80.0, 4000.0, 4)
        a = rng.uniform(0.2, 1.0, 4)
        phi = rng.uniform(0.0, 2 * np.pi, 4)
        env_f = rng.uniform(2.0, 5.0)
        x = np.zeros(n, np.float64)
        for k in range(4):
            x += a[k] * np.sin(2 * np.pi * f[k] * t + phi[k])
        env = 0.5 * (1.0 - np.cos(2 * np.pi * env_f * t))
        x = 0.3 * x * env + 0.05 * rng.standard_normal(n)
        out[i] = np.clip(x, -1.0, 1.0).astype(np.float32)
    return out


def frame(audio: np.ndarray, idx: int) -> np.ndarray:
    """``[B, 1920]`` slice holding frame ``idx`` of every stream."""
    return np.ascontiguousarray(audio[:, idx * FRAME_SIZE:(idx + 1) * FRAME_SIZE])
