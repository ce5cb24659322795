"""GPU parity of asr::State (step_pcm / step_tokens / reset_batch_idx) through the C ABI.

Bar: greedy token ids identical to the oracle's, logits within rel 1e-2, identical AsrMsg streams."""
import numpy as np
import pytest
import torch

from helpers import (GOLDEN, active, frame, golden_masks, lm_weights, mimi_weights, oracle_lm_weights, random_masks,
                     rel_err, synth_audio, tiny_lm, tiny_mimi)

pytestmark = pytest.mark.gpu


def _strip(msgs):
    return [(m[0], m[1]) if m[0] == "step" else m for m in msgs]


def _build(b, lspec, with_mimi=True):
    from dsm_b200 import api
    from oracle.lm import AsrState as OAsr, LmModel as OLm
    from oracle.mimi import Mimi as OMimi
    mspec = tiny_mimi()
    mw = mimi_weights(mspec, seed=0, layer_scale=0.5, decoder=False)
    lw = lm_weights(lspec, seed=0)
    gm = api.Mimi(mspec, mw, b, decoder=False) if with_mimi else None
    gl = api.LmModel(lspec, lw, b)
    ga = api.AsrState(gm, gl)
    om = OMimi(mspec, mw, b, decoder=False) if with_mimi else None
    oa = OAsr(b, lspec.asr_delay_in_tokens, om, OLm(lspec, oracle_lm_weights(lw), b))
    return ga, oa, (gm, gl)


@pytest.mark.parametrize("vocab", [8, 256])
def test_step_pcm_tokens_and_messages(vocab):
    b, frames = 3, 14
    lspec = tiny_lm(text_in_vocab_size=vocab + 1, text_out_vocab_size=vocab)
    ga, oa, _keep = _build(b, lspec)
    audio = synth_audio(b, frames)
    for i, m in enumerate(random_masks(frames, b, seed=11, first_active=3)):
        pcm = frame(audio, i)
        om = oa.step_pcm(torch.from_numpy(pcm).unsqueeze(1), m)
        gm = ga.step_pcm(pcm, m)
        a = active(m, b)
        assert rel_err(ga.last_logits()[a], oa.last_logits.numpy()[a]) < 1e-2
        assert (ga.last_tokens[a] == oa.last_tokens.numpy()[a]).all()
        assert _strip(gm) == _strip(om)
        for x, y in zip(gm, om):
            if x[0] == "step":
                assert np.abs(x[2][:, a] - y[2].numpy()[:, a]).max() < 2e-2      # VAD head probabilities
    assert ga.model_step_idx() == frames


def test_step_tokens_and_reset_batch_idx():
    b = 2
    lspec = tiny_lm(text_in_vocab_size=9, text_out_vocab_size=8)
    ga, oa, _keep = _build(b, lspec, with_mimi=False)
    rng = np.random.default_rng(2)
    for i in range(12):
        if i == 5:
            ga.reset_batch_idx(1)
            oa.batch[1] = type(oa.batch[1])(lspec.text_start_token)
            oa.lm.reset_batch_idx(1)
        codes = rng.integers(0, 2048, (b, lspec.audio_codebooks)).astype(np.uint32)
        om = oa.step_tokens(torch.from_numpy(codes.astype(np.int64)).unsqueeze(-1), None)
        gm = ga.step_tokens(codes, None)
        assert (ga.last_tokens == oa.last_tokens.numpy()).all()
        assert _strip(gm) == _strip(om)


def test_asr_matches_golden_vectors():
    gold = np.load(f"{GOLDEN}/asr_tiny.npz")
    lspec = tiny_lm(text_in_vocab_size=9, text_out_vocab_size=8)
    ga, _oa, _keep = _build(2, lspec)
    audio = synth_audio(2, len(golden_masks()))
    for i, m in enumerate(golden_masks()):
        msgs = ga.step_pcm(frame(audio, i), m)
        a = active(m, 2)
        assert (ga.last_tokens[a] == gold["tokens"][i][a]).all()
        assert rel_err(ga.last_logits()[a], gold["logits"][i][a]) < 1e-2
        assert repr(_strip(msgs)) == str(gold["msgs"][i])


def test_batch_size_mismatch_is_an_error():
    from dsm_b200 import api
    from dsm_b200._lib import DsmError
    mspec, lspec = tiny_mimi(), tiny_lm()
    gm = api.Mimi(mspec, mimi_weights(mspec, seed=0, decoder=False), 2, decoder=False)
    gl = api.LmModel(lspec, lm_weights(lspec, seed=0), 3)
    with pytest.raises(DsmError) as e:       # asr.rs:158-160 "batch size mismatch"
        api.AsrState(gm, gl)
    assert e.value.code == -2


def test_sm_partition_follows_the_measured_rule():
    """An ASR state with the bf16 ring moves its branches onto disjoint SM partitions (runtime.cu: sm_split_choice)
    and says so.  (DSM_SM_SPLIT overrides the rule: then only consistency is checked.)"""
    import os
    for B in (2, 8):
        ga, oa, _keep = _build(B, tiny_lm())
        lm_sms, codec_sms = ga.sm_partition()
        forced = os.environ.get("DSM_SM_SPLIT")
        if forced is None:
            assert (lm_sms, codec_sms) == (96, 52)
        elif int(forced) == 0:
            assert (lm_sms, codec_sms) == (0, 0)
        else:
            assert lm_sms >= int(forced) and codec_sms > 0 and lm_sms + codec_sms <= 148
        audio = synth_audio(B, 4)
        for i in range(4):
            pcm = frame(audio, i)
            om = oa.step_pcm(torch.from_numpy(pcm).unsqueeze(1), None)
            gm = ga.step_pcm(pcm, None)
            assert (ga.last_tokens == oa.last_tokens.numpy()).all()
            assert _strip(gm) == _strip(om)
