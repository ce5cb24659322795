"""Generates tests/golden/*.npz from the CPU oracle (PyTorch f32 restatement of the reference).

    python tests/golden/make_golden.py

The reference itself cannot run here (Rust + un-vendored candle, no cargo), so these vectors come from
the oracle in this container; they pin the oracle against accidental change (CPU test) and give the
GPU tests a fixed target that does not depend on the oracle's runtime.  Inputs: seeded random-init
weights (dsm_b200.weights, seed 0) and synthetic audio (dsm_b200.synth), both reproducible.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from dsm_b200.spec import tiny_lm, tiny_mimi
from dsm_b200.synth import frame, synth_audio
from dsm_b200.weights import lm_weights, mimi_weights, round_bf16
from oracle.lm import AsrState, LmModel
from oracle.mimi import Mimi

HERE = os.path.dirname(os.path.abspath(__file__))
B, FRAMES = 2, 8


def golden_masks():
    """Mask schedule shared by the generator and the tests (None = all active)."""
    return [None, None, [1, 1], [1, 0], [0, 1], [1, 1], [1, 0], [1, 1]]


def main():
    torch.set_num_threads(1)       # fixed summation order
    mspec, lspec = tiny_mimi(), tiny_lm(text_in_vocab_size=9, text_out_vocab_size=8)
    mw = mimi_weights(mspec, seed=0, layer_scale=0.5)
    lw = lm_weights(lspec, seed=0)
    lwr = {k: (round_bf16(v) if v.ndim == 2 else v) for k, v in lw.items()}
    audio = synth_audio(B, FRAMES)

    # ---- Mimi encode / decode ----
    enc = Mimi(mspec, mw, B, decoder=False)
    dec = Mimi(mspec, mw, B, decoder=True)
    codes, latents, pcms = [], [], []
    for i, m in enumerate(golden_masks()):
        c = enc.encode_step(torch.from_numpy(frame(audio, i)).unsqueeze(1), m)
        codes.append(c[:, :, 0].numpy().astype(np.uint32))
        latents.append(enc._last_latent.copy())
        p = dec.decode_step(c, m)
        pcms.append(p[:, 0].numpy().copy())
    np.savez_compressed(os.path.join(HERE, "mimi_tiny.npz"), codes=np.stack(codes), latents=np.stack(latents),
                        pcm=np.stack(pcms).astype(np.float32))

    # ---- ASR (Mimi + LM) ----
    asr = AsrState(B, lspec.asr_delay_in_tokens, Mimi(mspec, mw, B, decoder=False), LmModel(lspec, lwr, B))
    tokens, logits, msgs = [], [], []
    for i, m in enumerate(golden_masks()):
        out = asr.step_pcm(torch.from_numpy(frame(audio, i)).unsqueeze(1), m)
        tokens.append(asr.last_tokens.numpy().astype(np.uint32))
        logits.append(asr.last_logits.numpy().copy())
        msgs.append(repr([(x[0], x[1]) if x[0] == "step" else x for x in out]))
    np.savez_compressed(os.path.join(HERE, "asr_tiny.npz"), tokens=np.stack(tokens), logits=np.stack(logits),
                        msgs=np.array(msgs))
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
