"""GPU parity of Mimi encode_step / decode_step through the C ABI against the CPU oracle.

Bars (BASELINE.json north_star): RVQ code indices bit-identical on >= 99.5 % of frames (here: every
active frame of these cases), codec latents within rel 1e-2 (measured ~5e-6: the codec runs in true f32).
"""
import numpy as np
import pytest
import torch

from helpers import (GOLDEN, MimiSpec, active, frame, golden_masks, mimi_weights, random_masks, rel_err, synth_audio,
                     tiny_mimi)

pytestmark = pytest.mark.gpu


def _pair(spec, b, encoder=True, decoder=False, seed=0):
    from dsm_b200 import api
    from oracle.mimi import Mimi as OMimi
    w = mimi_weights(spec, seed=seed, layer_scale=0.5, decoder=decoder)
    return api.Mimi(spec, w, b, encoder=encoder, decoder=decoder), OMimi(spec, w, b, decoder=decoder)


@pytest.mark.parametrize("spec,b,frames", [(tiny_mimi(), 3, 12), (MimiSpec(), 2, 4)], ids=["ring6", "v0_1"])
def test_encode_step_matches_oracle(spec, b, frames):
    g, o = _pair(spec, b)
    audio = synth_audio(b, frames)
    same = total = 0
    for i, m in enumerate(random_masks(frames, b, seed=5)):
        pcm = frame(audio, i)
        oc = o.encode_step(torch.from_numpy(pcm).unsqueeze(1), m)[:, :, 0].numpy()
        gc = g.encode_step(pcm, m)
        a = active(m, b)
        assert rel_err(g.debug_tensor("seanet_out")[a], o._last_seanet[a]) < 1e-4
        assert rel_err(g.debug_tensor("transformer_out")[a], o._last_transformer[a]) < 1e-4
        assert rel_err(g.last_latent()[a], o._last_latent[a]) < 1e-4
        same += int((gc.astype(np.int64)[a] == oc[a]).all(axis=1).sum())
        total += int(a.sum())
    assert same >= 0.995 * total, f"only {same}/{total} frames have bit-identical RVQ codes"


def test_encode_matches_golden_vectors():
    from dsm_b200 import api
    gold = np.load(f"{GOLDEN}/mimi_tiny.npz")
    spec = tiny_mimi()
    g = api.Mimi(spec, mimi_weights(spec, seed=0, layer_scale=0.5), 2, decoder=True)
    audio = synth_audio(2, len(golden_masks()))
    for i, m in enumerate(golden_masks()):
        codes = g.encode_step(frame(audio, i), m)
        a = active(m, 2)
        assert (codes[a] == gold["codes"][i][a]).all()
        assert rel_err(g.last_latent()[a], gold["latents"][i][a]) < 1e-4
        pcm = g.decode_step(gold["codes"][i], m)
        assert rel_err(pcm[a], gold["pcm"][i][a]) < 1e-4


@pytest.mark.parametrize("spec,b,frames", [(tiny_mimi(), 3, 10), (MimiSpec(), 2, 3)], ids=["ring6", "v0_1"])
def test_decode_step_matches_oracle(spec, b, frames):
    g, o = _pair(spec, b, encoder=False, decoder=True)
    rng = np.random.default_rng(7)
    for i, m in enumerate(random_masks(frames, b, seed=9)):
        codes = rng.integers(0, spec.bins, (b, spec.n_q)).astype(np.uint32)
        op = o.decode_step(torch.from_numpy(codes.astype(np.int64)).unsqueeze(-1), m)[:, 0].numpy()
        gp = g.decode_step(codes, m)
        a = active(m, b)
        assert rel_err(gp[a], op[a]) < 1e-4


def test_reset_batch_idx_and_reset_state():
    spec, b = tiny_mimi(), 2
    g, o = _pair(spec, b)
    audio = synth_audio(b, 8)
    for i in range(3):
        g.encode_step(frame(audio, i))
        o.encode_step(torch.from_numpy(frame(audio, i)).unsqueeze(1))
    g.reset_batch_idx(1)
    o.reset_batch_idx(1)
    for i in range(3, 6):
        gc = g.encode_step(frame(audio, i))
        oc = o.encode_step(torch.from_numpy(frame(audio, i)).unsqueeze(1))[:, :, 0].numpy()
        assert (gc.astype(np.int64) == oc).all()
    g.reset_state()
    o.reset_state()
    for i in range(2):
        gc = g.encode_step(frame(audio, i))
        oc = o.encode_step(torch.from_numpy(frame(audio, i)).unsqueeze(1))[:, :, 0].numpy()
        assert (gc.astype(np.int64) == oc).all()


def test_device_pointers_and_graph_vs_eager():
    """Device-resident PCM in / codes out give the same codes as host buffers; a CUDA-graph replay gives
    the same codes as eager launches."""
    from dsm_b200 import api
    spec, b = tiny_mimi(), 2
    w = mimi_weights(spec, seed=0, layer_scale=0.5, decoder=False)
    audio = synth_audio(b, 4)
    host = api.Mimi(spec, w, b, decoder=False)
    ref = [host.encode_step(frame(audio, i)).copy() for i in range(4)]
    dev = api.Mimi(spec, w, b, decoder=False)
    out = torch.empty((b, spec.n_q), dtype=torch.int32, device="cuda")
    for i in range(4):
        dev.encode_step(torch.from_numpy(frame(audio, i)).cuda(), None, out=out)
        dev.synchronize()
        assert (out.cpu().numpy().astype(np.uint32) == ref[i]).all()
    api.set_graphs(False)
    try:
        eager = api.Mimi(spec, w, b, decoder=False)
        for i in range(4):
            assert (eager.encode_step(frame(audio, i)) == ref[i]).all()
    finally:
        api.set_graphs(True)


def test_shape_errors_are_reported_not_fatal():
    from dsm_b200 import api
    from dsm_b200._lib import DsmError
    spec = tiny_mimi()
    g = api.Mimi(spec, mimi_weights(spec, seed=0, decoder=False), 2, decoder=False)
    with pytest.raises(DsmError):
        g.decode_step(np.zeros((2, spec.n_q), np.uint32))       # no decoder on this handle
    with pytest.raises(DsmError):
        g.reset_batch_idx(5)
    w = mimi_weights(spec, seed=0, decoder=False)
    del w["downsample.conv.conv.conv.weight"]
    with pytest.raises(DsmError) as e:
        api.Mimi(spec, w, 2, decoder=False)
    assert e.value.code == -6


def test_v0_1_codec_ring_full_and_wrapped():
    """Mimi v0_1 at its real widths with the 250-slot codec rings (encoder AND decoder transformer) full and wrapped:
    135 frames = 270 positions.  Codes of every frame, and the decoded audio of every frame, against the oracle."""
    import os
    torch.set_num_threads(os.cpu_count() or 8)
    spec, b, frames = MimiSpec(), 2, 135
    g, o = _pair(spec, b, encoder=True, decoder=True)
    audio = synth_audio(b, frames)
    rng = np.random.default_rng(3)
    same = 0
    worst_lat = worst_pcm = 0.0
    for i in range(frames):
        m = None if i % 50 != 49 else np.array([1, 0], bool)            # a few masked steps on the way
        pcm = frame(audio, i)
        oc = o.encode_step(torch.from_numpy(pcm).unsqueeze(1), m)[:, :, 0].numpy()
        gc = g.encode_step(pcm, m)
        a = active(m, b)
        same += int((gc.astype(np.int64)[a] == oc[a]).all(axis=1).sum())
        if i >= 125:
            worst_lat = max(worst_lat, rel_err(g.last_latent()[a], o._last_latent[a]))
        codes = rng.integers(0, spec.bins, (b, spec.n_q)).astype(np.uint32)
        op = o.decode_step(torch.from_numpy(codes.astype(np.int64)).unsqueeze(-1), m)[:, 0].numpy()
        gp = g.decode_step(codes, m)
        if i >= 125:
            worst_pcm = max(worst_pcm, rel_err(gp[a], op[a]))
    total = frames * b - len(range(49, frames, 50))
    print(f"v0_1 ring 250 wrapped: codes identical {same}/{total}, latent rel {worst_lat:.2e}, pcm rel {worst_pcm:.2e}")
    assert same >= 0.995 * total
    assert worst_lat < 1e-4 and worst_pcm < 1e-4
