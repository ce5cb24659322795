"""Stage-by-stage parity report of the CUDA path against the oracle (run on the GPU box).

    python tools/gpu_check.py [mimi|lm|asr|decode|all]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from dsm_b200 import api
from dsm_b200.spec import MimiSpec, tiny_lm, tiny_mimi, stt_1b_en_fr
from dsm_b200.synth import frame, synth_audio
from dsm_b200.weights import lm_weights, mimi_weights, round_bf16


def err(name, got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    d = np.abs(got - ref).max()
    s = np.abs(ref).max()
    print(f"  {name:24s} max|d|={d:.3e}  max|ref|={s:.3e}  rel={d / (s + 1e-30):.3e}", flush=True)
    return d / (s + 1e-30)


def check_mimi(spec, B=3, frames=10, masks=True):
    from oracle.mimi import Mimi as OMimi
    print(f"== mimi encode  ctx={spec.context} layers={spec.num_layers} B={B}")
    w = mimi_weights(spec, seed=0, layer_scale=0.5)
    om = OMimi(spec, w, B, decoder=False)
    gm = api.Mimi(spec, w, B, decoder=False)
    audio = synth_audio(B, frames)
    rng = np.random.default_rng(5)
    same = total = 0
    for i in range(frames):
        pcm = frame(audio, i)
        mask = None
        if masks and i >= 2:
            mask = (rng.random(B) > 0.3)
            if not mask.any():
                mask[0] = True
        ocodes = om.encode_step(torch.from_numpy(pcm).unsqueeze(1), mask)
        gcodes = gm.encode_step(pcm, mask)
        act = np.ones(B, bool) if mask is None else mask
        print(f" frame {i} mask={None if mask is None else mask.astype(int).tolist()}")
        # oracle intermediate taps
        err("seanet_out", gm.debug_tensor("seanet_out")[act], om._last_seanet[act])
        err("transformer_out", gm.debug_tensor("transformer_out")[act], om._last_transformer[act])
        err("latent", gm.last_latent()[act], om._last_latent[act])
        oc = ocodes[:, :, 0].numpy().astype(np.int64)
        eq = (oc == gcodes.astype(np.int64))
        same += int(eq[act].all(axis=1).sum())
        total += int(act.sum())
        print(f"  codes identical rows {int(eq[act].all(axis=1).sum())}/{int(act.sum())}; first diff cb:",
              [int(np.argmin(r)) if not r.all() else -1 for r in eq[act]])
    print(f" frames with identical codes: {same}/{total}")
    gm.close()


def check_lm(spec, B=3, steps=16):
    from oracle.lm import LmModel as OLm
    print(f"== lm forward {spec.name} B={B}")
    w = lm_weights(spec, seed=0)
    wr = {k: (round_bf16(v) if v.ndim == 2 else v) for k, v in w.items()}
    ol = OLm(spec, wr, B)
    gl = api.LmModel(spec, w, B)
    rng = np.random.default_rng(3)
    for i in range(steps):
        text = rng.integers(0, spec.text_in_vocab_size, B)
        audio = rng.integers(0, spec.audio_vocab_size, (B, spec.audio_codebooks))
        mask = None
        if i >= 3:
            mask = rng.random(B) > 0.3
            if not mask.any():
                mask[0] = True
        act = np.ones(B, bool) if mask is None else mask
        ologits, oys = ol.forward_cond(torch.from_numpy(text), torch.from_numpy(audio), None,
                                       [True] * B if mask is None else mask.tolist())
        glogits, gys = gl.forward_cond(text, audio, None, mask)
        print(f" step {i} mask={None if mask is None else mask.astype(int).tolist()}")
        err("ys", gys[act], oys[:, 0].numpy()[act])
        err("logits", glogits[act], ologits[:, 0].numpy()[act])
        print("  argmax equal:", (glogits.argmax(-1) == ologits[:, 0].numpy().argmax(-1))[act].tolist())
    gl.close()


def check_decode(spec, B=2, frames=6):
    from oracle.mimi import Mimi as OMimi
    print(f"== mimi decode ctx={spec.context} B={B}")
    w = mimi_weights(spec, seed=0, layer_scale=0.5)
    om = OMimi(spec, w, B, decoder=True)
    gm = api.Mimi(spec, w, B, encoder=False, decoder=True)
    rng = np.random.default_rng(7)
    for i in range(frames):
        codes = rng.integers(0, spec.bins, (B, spec.n_q)).astype(np.uint32)
        mask = None
        if i >= 2:
            mask = rng.random(B) > 0.3
            if not mask.any():
                mask[0] = True
        act = np.ones(B, bool) if mask is None else mask
        opcm = om.decode_step(torch.from_numpy(codes.astype(np.int64)).unsqueeze(-1), mask)
        gpcm = gm.decode_step(codes, mask)
        print(f" frame {i}")
        err("pcm", gpcm[act], opcm[:, 0].numpy()[act])
    gm.close()


def check_asr(mspec, lspec, B=3, frames=12):
    from oracle.lm import AsrState as OAsr, LmModel as OLm
    from oracle.mimi import Mimi as OMimi
    print(f"== asr step_pcm B={B}")
    mw = mimi_weights(mspec, seed=0, layer_scale=0.5, decoder=False)
    lw = lm_weights(lspec, seed=0)
    lwr = {k: (round_bf16(v) if v.ndim == 2 else v) for k, v in lw.items()}
    om = OMimi(mspec, mw, B, decoder=False)
    ol = OLm(lspec, lwr, B)
    oa = OAsr(B, lspec.asr_delay_in_tokens, om, ol)
    gm = api.Mimi(mspec, mw, B, decoder=False)
    gl = api.LmModel(lspec, lw, B)
    ga = api.AsrState(gm, gl)
    audio = synth_audio(B, frames)
    rng = np.random.default_rng(11)
    for i in range(frames):
        pcm = frame(audio, i)
        mask = None
        if i >= 3:
            mask = rng.random(B) > 0.3
            if not mask.any():
                mask[0] = True
        act = np.ones(B, bool) if mask is None else mask
        omsgs = oa.step_pcm(torch.from_numpy(pcm).unsqueeze(1), mask)
        gmsgs = ga.step_pcm(pcm, mask)
        print(f" frame {i}: tokens gpu={ga.last_tokens[act].tolist()} oracle={oa.last_tokens.numpy()[act].tolist()}")
        err("logits", ga.last_logits()[act], oa.last_logits.numpy()[act])
        print("  msgs gpu   :", [m if m[0] != 'step' else ('step', m[1]) for m in gmsgs])
        print("  msgs oracle:", [m if m[0] != 'step' else ('step', m[1]) for m in omsgs])
    ga.close(); gl.close(); gm.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.set_num_threads(os.cpu_count() or 8)
    t0 = time.time()
    if what in ("mimi", "all"):
        check_mimi(tiny_mimi())
        check_mimi(MimiSpec(), B=2, frames=4, masks=False)
    if what in ("lm", "all"):
        check_lm(tiny_lm())
    if what in ("decode", "all"):
        check_decode(tiny_mimi())
    if what in ("asr", "all"):
        check_asr(tiny_mimi(), tiny_lm())
    print(f"done in {time.time() - t0:.1f}s")
