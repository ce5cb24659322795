"""The scheduler drives the same ASR state as manual step_pcm calls: streams that join late, pause and
leave produce, per connection, exactly the tokens/messages of an AsrState fed the same frames and mask."""
import numpy as np
import pytest

from helpers import frame, lm_weights, mimi_weights, synth_audio, tiny_lm, tiny_mimi

pytestmark = pytest.mark.gpu


def test_scheduler_matches_manual_stepping():
    from dsm_b200 import api
    B, frames = 3, 10
    mspec, lspec = tiny_mimi(), tiny_lm(text_in_vocab_size=9, text_out_vocab_size=8)
    mw, lw = mimi_weights(mspec, seed=0, layer_scale=0.5, decoder=False), lm_weights(lspec, seed=0)

    def build():
        m = api.Mimi(mspec, mw, B, decoder=False)
        l = api.LmModel(lspec, lw, B)
        return api.AsrState(m, l), m, l

    asr_s, *keep_s = build()
    sched = api.BatchedAsr(asr_s)
    asr_m, *keep_m = build()
    audio = synth_audio(B, frames)

    # slot 0 streams from tick 0, slot 1 joins at tick 2, slot 0 pauses at ticks 4-5
    s0, _ = sched.open()
    s1 = None
    cursor = {0: 0, 1: 0}
    for tick in range(frames):
        if tick == 2:
            s1, _ = sched.open()
        feed = {s0: not (4 <= tick <= 5)}
        if s1 is not None:
            feed[s1] = True
        mask = np.zeros(B, np.uint8)
        pcm = np.zeros((B, 1920), np.float32)
        for slot, on in feed.items():
            if on:
                fr = frame(audio, cursor[slot])[slot]
                sched.push_audio(slot, fr[:700])              # arbitrary packetisation
                sched.push_audio(slot, fr[700:])
                pcm[slot], mask[slot] = fr, 1
                cursor[slot] += 1
        if tick == 2:
            asr_m.reset_batch_idx(s1)                         # what InMsg::Init does for a new connection
        if tick == 0:
            asr_m.reset_batch_idx(s0)
        stepped, outs = sched.tick()
        assert stepped == bool(mask.any())
        assert sched.last_mask().tolist() == mask.tolist()
        manual = asr_m.step_pcm(pcm, mask)
        act = mask.astype(bool)
        if act.any():
            # the scheduler steps the C++ state directly; read its sampled tokens back through the logits
            assert (asr_s.last_logits().argmax(-1)[act] == asr_m.last_tokens[act]).all()
        words_s = sorted((o["slot"], tuple(o["tokens"]), o["start_time"]) for o in outs if o["kind"] == "word")
        words_m = sorted((m[3], tuple(m[1]), m[2]) for m in manual if m[0] == "word")
        assert words_s == words_m
        steps = [o for o in outs if o["kind"] == "step"]
        assert sorted(o["slot"] for o in steps) == [i for i in range(B) if mask[i]]
        assert all(len(o["prs"]) == lspec.extra_heads_num for o in steps)


def test_threaded_router_matches_manual_stepping():
    """Pipelined tick thread (step_begin / step_end, double-buffered pinned batch) + event queues: whatever way the
    frames end up batched, every connection receives exactly the words of an AsrState fed its frames."""
    import time
    from dsm_b200 import api
    B, frames = 3, 14
    mspec, lspec = tiny_mimi(), tiny_lm(text_in_vocab_size=9, text_out_vocab_size=8)
    mw, lw = mimi_weights(mspec, seed=0, layer_scale=0.5, decoder=False), lm_weights(lspec, seed=0)

    def build():
        m = api.Mimi(mspec, mw, B, decoder=False)
        l = api.LmModel(lspec, lw, B)
        return api.AsrState(m, l), m, l

    asr_s, *keep_s = build()
    sched = api.BatchedAsr(asr_s)
    asr_m, *keep_m = build()
    audio = synth_audio(B, frames)
    conns = [sched.open() for _ in range(2)]
    sched.start()
    try:
        for f in range(frames):
            for slot, cid in conns:
                sched.push_audio(slot, frame(audio, f)[slot], channel_id=cid)
            if f % 3 == 0:
                time.sleep(0.01)                               # let the ticks interleave with the pushes
        deadline = time.time() + 30
        while time.time() < deadline and sched.stats(0)["frames"] < frames * len(conns):
            time.sleep(0.005)
        time.sleep(0.05)
    finally:
        sched.stop()
    assert sched.stats(0)["frames"] == frames * len(conns)
    got = {slot: [] for slot, _ in conns}
    for slot, _ in conns:
        while True:
            e = sched.poll(slot)
            if e is None:
                break
            if e["kind"] == "word":
                got[slot].append((tuple(e["tokens"]), e["start_time"]))
    want = {slot: [] for slot, _ in conns}
    mask = np.zeros(B, np.uint8)
    for slot, _ in conns:
        mask[slot] = 1
        asr_m.reset_batch_idx(slot)
    for f in range(frames):
        for m in asr_m.step_pcm(frame(audio, f), mask):
            if m[0] == "word":
                want[m[3]].append((tuple(m[1]), m[2]))
    assert got == want


def test_router_over_two_devices():
    """Handles on device 0 AND device 1 in one process (per-device kernel attributes), one scheduler over both:
    connections are spread over the GPUs and every connection gets the words of a single-device manual run."""
    import time
    from dsm_b200 import _lib, api
    if _lib.load().dsm_device_count() < 2:
        pytest.skip("needs two GPUs")
    B, frames = 3, 12
    mspec, lspec = tiny_mimi(), tiny_lm(text_in_vocab_size=9, text_out_vocab_size=8)
    mw, lw = mimi_weights(mspec, seed=0, layer_scale=0.5, decoder=False), lm_weights(lspec, seed=0)

    def build(dev):
        m = api.Mimi(mspec, mw, B, device=dev, decoder=False)
        l = api.LmModel(lspec, lw, B, device=dev)
        return api.AsrState(m, l), m, l

    states = [build(0), build(1)]
    sched = api.BatchedAsr([s[0] for s in states])
    assert sched.n_devices == 2
    conns = [sched.open() for _ in range(4)]
    assert sorted(slot // B for slot, _ in conns) == [0, 0, 1, 1]          # least-loaded placement
    audio = synth_audio(4, frames)
    sched.start()
    try:
        for f in range(frames):
            for i, (slot, cid) in enumerate(conns):
                sched.push_audio(slot, frame(audio, f)[i], channel_id=cid)
        deadline = time.time() + 60
        while time.time() < deadline and sum(sched.stats(g)["frames"] for g in range(2)) < frames * len(conns):
            time.sleep(0.005)
        time.sleep(0.05)
    finally:
        sched.stop()
    got = []
    for slot, _ in conns:
        words = []
        while True:
            e = sched.poll(slot)
            if e is None:
                break
            if e["kind"] == "word":
                words.append((tuple(e["tokens"]), e["start_time"]))
        got.append(words)
    ref, *keep = build(0)
    want = [[] for _ in conns]
    padded = np.zeros((frames, B, 1920), np.float32)
    for rnd in range(0, len(conns), B):                                      # B connections at a time through one state
        ref.reset()
        group = list(range(rnd, min(rnd + B, len(conns))))
        mask = np.zeros(B, np.uint8)
        mask[:len(group)] = 1
        for f in range(frames):
            pcm = np.zeros((B, 1920), np.float32)
            for j, i in enumerate(group):
                pcm[j] = frame(audio, f)[i]
            for m in ref.step_pcm(pcm, mask):
                if m[0] == "word":
                    want[group[m[3]]].append((tuple(m[1]), m[2]))
    assert got == want
