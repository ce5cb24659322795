"""A few full ASR steps at the bench workload, for ncu launch lists / --set full captures.

    python tools/profile_steps.py [--streams 256] [--steps 3] [--eager]

The LM ring position is fast-forwarded (dsm_lm_debug_set_position) so the attention kernel reads full
rings without 750 untimed steps; numbers printed by this script are not benchmark values.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from dsm_b200 import api
from dsm_b200.spec import MimiSpec, stt_1b_en_fr, stt_2_6b_en
from dsm_b200.synth import frame, synth_audio

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=256)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--eager", action="store_true")
ap.add_argument("--model", default="1b")
ap.add_argument("--workload", default="asr", choices=["asr", "mimi", "decode", "lm"])
ap.add_argument("--kv", default="bf16", choices=["bf16", "int8", "int8k"])
a = ap.parse_args()
if a.eager:
    api.set_graphs(False)
B = a.streams
mspec = MimiSpec()
lspec = stt_1b_en_fr() if a.model == "1b" else stt_2_6b_en()
audio = synth_audio(min(B, 8), a.steps)
audio = np.tile(audio, (B // audio.shape[0] + 1, 1))[:B]
if a.workload == "asr":
    mimi = api.Mimi(mspec, None, B, decoder=False)
    lm = api.LmModel(lspec, None, B, kv_dtype=a.kv)
    asr = api.AsrState(mimi, lm)
    lm.debug_set_position(2000)
    for i in range(a.steps):
        asr.step_pcm(frame(audio, i), None, want_messages=False)
        print("step", i, "ms", asr.last_step_ms())
elif a.workload == "lm":
    import time
    lm = api.LmModel(lspec, None, B, kv_dtype=a.kv)
    asr = api.AsrState(None, lm)
    lm.debug_set_position(2000)
    codes = np.random.default_rng(7).integers(0, 2048, (B, 32)).astype(np.uint32)
    for i in range(a.steps):
        t0 = time.perf_counter()
        asr.step_tokens(codes, None, want_messages=False)
        print("lm step", i, "wall ms", (time.perf_counter() - t0) * 1e3, "device ms", asr.last_step_ms()[0])
    t0 = time.perf_counter()
    for i in range(50):
        asr.step_tokens(codes, None, want_messages=False)
    print("lm branch alone: %.3f ms/step over 50 steps" % ((time.perf_counter() - t0) * 1e3 / 50))
elif a.workload == "mimi":
    import time
    import torch
    mimi = api.Mimi(mspec, None, B, decoder=False)
    pcm = [torch.from_numpy(frame(audio, i)).cuda() for i in range(a.steps)]
    out = torch.empty((B, 32), dtype=torch.int32, device="cuda")
    for i in range(a.steps):
        mimi.synchronize()
        t0 = time.perf_counter()
        mimi.encode_step(pcm[i], None, out=out)
        mimi.synchronize()
        print("mimi step", i, "wall ms", (time.perf_counter() - t0) * 1e3)
    mimi.synchronize()
    t0 = time.perf_counter()
    for i in range(50):
        mimi.encode_step(pcm[i % a.steps], None, out=out)
    mimi.synchronize()
    print("mimi branch alone: %.3f ms/step over 50 steps" % ((time.perf_counter() - t0) * 1e3 / 50))
else:
    mimi = api.Mimi(mspec, None, B, encoder=False, decoder=True)
    codes = np.random.default_rng(7).integers(0, 2048, (B, 32)).astype(np.uint32)
    for i in range(a.steps):
        mimi.decode_step(codes, None)
    print("ok")
