"""Size-independent properties at BASELINE.json's full size (stt-1b-en_fr, 256 streams, random init):
determinism, stream independence (permuting the batch permutes the outputs), masked streams leave no
trace.  No oracle involved: the CPU oracle cannot run this size in seconds."""
import numpy as np
import pytest

from helpers import frame, synth_audio

pytestmark = pytest.mark.gpu

B = 256


def _asr():
    from dsm_b200 import api
    from dsm_b200.spec import MimiSpec, stt_1b_en_fr
    m = api.Mimi(MimiSpec(), None, B, decoder=False)
    l = api.LmModel(stt_1b_en_fr(), None, B)
    return api.AsrState(m, l), m, l


def test_permutation_equivariance_and_mask_isolation():
    audio = synth_audio(8, 6)
    audio = np.tile(audio, (B // 8, 1))
    audio[::3] *= 0.5                                   # make streams distinct in more than 8 ways
    rng = np.random.default_rng(0)
    perm = rng.permutation(B)
    a1, m1, l1 = _asr()
    toks1, codes1, logits1 = [], [], []
    from dsm_b200 import api
    from dsm_b200.spec import MimiSpec
    probe = api.Mimi(MimiSpec(), None, B, decoder=False)
    for i in range(6):
        a1.step_pcm(frame(audio, i))
        toks1.append(a1.last_tokens.copy())
        logits1.append(a1.last_logits())
        codes1.append(probe.encode_step(frame(audio, i)).copy())
    del a1, m1, l1
    a2, m2, l2 = _asr()
    probe2 = api.Mimi(MimiSpec(), None, B, decoder=False)
    agree = np.ones(B, bool)        # streams whose sampled history still matches (a flipped token changes the next input)
    for i in range(6):
        a2.step_pcm(frame(audio[perm], i))
        c2 = probe2.encode_step(frame(audio[perm], i))
        assert (c2 == codes1[i][perm]).all(), "RVQ codes depend on the position of a stream in the batch"
        # The LM is not bit-reproducible across batch layouts: split-K partial sums are combined with f32
        # atomics, and a 1e-7 difference can flip a bf16 rounding that then propagates.  Logits must
        # agree to bf16 tolerance; greedy tokens of near-uniform random-init logits may flip rarely.
        l2 = a2.last_logits()
        err = np.abs(l2 - logits1[i][perm])[agree].max() / np.abs(l2).max()
        assert err < 2e-2, f"logits depend on the position of a stream in the batch (rel {err:.3e})"
        agree &= a2.last_tokens == toks1[i][perm]
    assert agree.mean() >= 0.9, f"greedy tokens diverged on {100 * (1 - agree.mean()):.1f}% of streams"


def test_masked_streams_do_not_advance():
    a, m, l = _asr()
    audio = np.tile(synth_audio(8, 5), (B // 8, 1))
    mask = np.ones(B, np.uint8)
    mask[1::2] = 0
    ref, _, _ = _asr()
    for i in range(2):
        a.step_pcm(frame(audio, i))
        ref.step_pcm(frame(audio, i))
    a.step_pcm(frame(audio, 4), mask)                   # odd streams sit this frame out
    a.step_pcm(frame(audio, 2), 1 - mask)               # ... and the even ones the next
    ref.step_pcm(frame(audio, 2))
    odd = np.arange(B)[1::2]
    err = np.abs(a.last_logits()[odd] - ref.last_logits()[odd]).max() / np.abs(ref.last_logits()).max()
    assert err < 2e-2
    assert (a.last_tokens[odd] == ref.last_tokens[odd]).mean() >= 0.95
