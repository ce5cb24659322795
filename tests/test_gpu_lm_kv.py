"""LM KV ring formats at real widths with a FULL, WRAPPED 750-slot ring: bf16 (the reference's, default), int8
(K and V, per-slot scales) and int8k (K int8, V bf16) against the f32 oracle -- logits error and greedy-token
identity, as numbers (printed with -s, asserted against each format's bar) -- and run-to-run bit reproducibility of
the logits (split-K partials meet in cluster shared memory in rank order, not through atomics)."""
import os

import numpy as np
import pytest
import torch

from helpers import LmSpec, lm_weights, oracle_lm_weights, rel_err

pytestmark = pytest.mark.gpu

# logits bars (rel to max |logit|): north_star's bf16 tolerance is 1e-2.  Measured on B200 at these shapes: see
# DESIGN.md section 4; the int8 V ring sits at the tolerance, which is why it is opt-in.
BARS = {"bf16": 1e-2, "int8k": 1e-2, "int8": 2e-2}


@pytest.fixture(scope="module")
def wrapped_ring_run():
    """2 layers at stt-1b widths, ring 750, 4 streams, 790 steps (full after 750, wrapped after); every copy is fed
    the oracle's tokens so histories stay comparable; logits compared on every step."""
    from dsm_b200 import api
    from oracle.lm import LmModel as OLm
    torch.set_num_threads(os.cpu_count() or 8)
    spec = LmSpec(name="stt-1b-2layers", num_layers=2)          # context 750, 16 heads x 128
    b, steps = 4, 790
    w = lm_weights(spec, seed=0)
    o = OLm(spec, oracle_lm_weights(w), b)
    gs = {k: api.LmModel(spec, w, b, kv_dtype=k) for k in BARS}
    rng = np.random.default_rng(11)
    text = np.full(b, spec.text_start_token, np.int64)
    worst = {k: 0.0 for k in BARS}
    worst_wrapped = {k: 0.0 for k in BARS}
    same = {k: 0 for k in BARS}
    far = {k: 0 for k in BARS}          # token mismatches that are NOT near-ties of the oracle's logits
    for i in range(steps):
        audio = rng.integers(0, 2048, (b, spec.audio_codebooks))
        mask = None if i % 97 else np.array([1, 1, 0, 1], np.uint8)     # stream 2 sits a few steps out
        ol, _ = o.forward_cond(torch.from_numpy(text), torch.from_numpy(audio), None,
                               [True] * b if mask is None else [bool(x) for x in mask])
        ol = ol[:, 0].numpy()
        act = np.ones(b, bool) if mask is None else mask.astype(bool)
        o_tok = ol.argmax(-1)
        for k, g in gs.items():
            gl, _ = g.forward_cond(text, audio, None, mask, want_ys=False)
            e = rel_err(gl[act], ol[act])
            worst[k] = max(worst[k], e)
            if i >= spec.context:
                worst_wrapped[k] = max(worst_wrapped[k], e)
            g_tok = gl.argmax(-1)
            same[k] += int((g_tok[act] == o_tok[act]).sum())
            # a differing token must be one the oracle itself ranks within the logits tolerance of its argmax
            gap = ol[np.arange(b), o_tok] - ol[np.arange(b), g_tok]
            far[k] += int(((gap > BARS[k] * np.abs(ol).max()) & act).sum())
        text = np.where(act, o_tok, text)
    total = steps * b - len(range(0, steps, 97))
    for k in BARS:
        print(f"kv {k:6s}: logits rel err max {worst[k]:.3e} (wrapped ring {worst_wrapped[k]:.3e}), "
              f"greedy tokens identical {same[k]}/{total}, mismatches outside the tolerance {far[k]}")
    return worst, same, far, total


@pytest.mark.parametrize("kv", list(BARS))
def test_full_wrapped_ring_against_oracle(wrapped_ring_run, kv):
    worst, same, far, total = wrapped_ring_run
    assert worst[kv] < BARS[kv], f"{kv}: logits rel err {worst[kv]:.3e}"
    # Random-init logits over 8000 entries are nearly flat (top-2 gaps of 1e-3 of the range are common), so ~1 % of the
    # greedy tokens are near-ties that bf16 operands resolve either way; none may differ by more than the tolerance.
    assert far[kv] == 0, f"{kv}: {far[kv]} greedy tokens differ beyond the logits tolerance"
    assert same[kv] >= 0.98 * total, f"{kv}: tokens identical on {same[kv]}/{total}"


@pytest.mark.parametrize("kv", list(BARS))
def test_logits_bit_reproducible(kv):
    """Two handles, same inputs, 256 streams (every split-K shape of the step is in play): identical bits."""
    from dsm_b200 import api
    spec = LmSpec(name="stt-1b-2layers", num_layers=2, context=24)
    b = 256
    outs = []
    for rep in range(2):
        g = api.LmModel(spec, None, b, kv_dtype=kv)
        rng = np.random.default_rng(3)
        logits = []
        for i in range(30):
            text = rng.integers(0, spec.text_in_vocab_size, b)
            audio = rng.integers(0, 2048, (b, spec.audio_codebooks))
            logits.append(g.forward_cond(text, audio, None, None, want_ys=False)[0].copy())
        outs.append(np.stack(logits))
        del g
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


def test_stt_2_6b_head_dim_64_int8():
    """int8 ring at the stt-2.6b head shape (32 heads x 64, ring 375 -> padded to 376 slots), wrapped."""
    from dsm_b200 import api
    from oracle.lm import LmModel as OLm
    spec = LmSpec(name="stt-2.6b-2layers", num_heads=32, num_layers=2, context=375, text_in_vocab_size=4001,
                  text_out_vocab_size=4000, extra_heads_num=0, extra_heads_dim=0)
    b = 3
    w = lm_weights(spec, seed=0)
    o = OLm(spec, oracle_lm_weights(w), b)
    gs = {k: api.LmModel(spec, w, b, kv_dtype=k) for k in ("int8", "int8k")}
    rng = np.random.default_rng(5)
    worst = {k: 0.0 for k in gs}
    for i in range(400):
        text = rng.integers(0, spec.text_in_vocab_size, b)
        audio = rng.integers(0, 2048, (b, spec.audio_codebooks))
        ol, _ = o.forward_cond(torch.from_numpy(text), torch.from_numpy(audio), None, [True] * b)
        for k, g in gs.items():
            gl, _ = g.forward_cond(text, audio, None, None, want_ys=False)
            worst[k] = max(worst[k], rel_err(gl, ol[:, 0].numpy()))
    print({k: f"{v:.3e}" for k, v in worst.items()})
    assert worst["int8k"] < 1e-2 and worst["int8"] < 2e-2
