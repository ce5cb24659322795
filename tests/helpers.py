"""Shared fixtures for the parity tests: specs, seeded weights, synthetic audio, mask schedules."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from dsm_b200.spec import LmSpec, MimiSpec, tiny_lm, tiny_mimi  # noqa: E402
from dsm_b200.synth import frame, synth_audio  # noqa: E402
from dsm_b200.weights import lm_weights, mimi_weights, round_bf16  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def has_gpu() -> bool:
    return torch.cuda.is_available()


def oracle_lm_weights(w):
    """The device rounds matrices to bf16 on load; hand the oracle the same rounded values so weight
    rounding is not part of the comparison (norm scales stay f32 on both sides)."""
    return {k: (round_bf16(v) if v.ndim == 2 else v) for k, v in w.items()}


def random_masks(n_steps, b, seed, p_inactive=0.3, first_active=2):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_steps):
        if i < first_active:
            out.append(None)
            continue
        m = rng.random(b) > p_inactive
        if not m.any():
            m[0] = True
        out.append(m)
    return out


def active(mask, b):
    return np.ones(b, bool) if mask is None else np.asarray(mask, bool)


def rel_err(got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.abs(got - ref).max() / (np.abs(ref).max() + 1e-30))


def golden_masks():
    return [None, None, [1, 1], [1, 0], [0, 1], [1, 1], [1, 0], [1, 1]]
