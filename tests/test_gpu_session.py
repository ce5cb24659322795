"""The non-batched `Asr` path (moshi-server/src/asr.rs:95-330) as a single-connection session over a batch-1
asr::State: wire bytes in (Audio messages of arbitrary length), wire bytes out with SentencePiece text, token log as
safetensors -- against the same frames stepped by hand."""
import ctypes as C
import json
import os
import struct

import msgpack
import numpy as np
import pytest

from helpers import GOLDEN, frame, lm_weights, mimi_weights, synth_audio, tiny_lm, tiny_mimi

pytestmark = pytest.mark.gpu


def test_session_matches_manual_stepping(tmp_path):
    from dsm_b200 import _lib, api
    from dsm_b200._lib import check
    lib = _lib.load()
    frames = 16
    mspec = tiny_mimi()
    lspec = tiny_lm(text_in_vocab_size=601, text_out_vocab_size=600)      # token ids inside the fixture tokenizer
    mw, lw = mimi_weights(mspec, seed=0, layer_scale=0.5, decoder=False), lm_weights(lspec, seed=0)

    def build():
        m = api.Mimi(mspec, mw, 1, decoder=False)
        l = api.LmModel(lspec, lw, 1)
        return api.AsrState(m, l), m, l

    asr_s, *keep_s = build()
    asr_m, *keep_m = build()
    spm = C.c_void_p()
    check(lib.dsm_spm_open(os.path.join(GOLDEN, "tiny_spm.model").encode(), C.byref(spm)))
    sess = C.c_void_p()
    check(lib.dsm_session_create(asr_s._h, spm, C.byref(sess)))
    audio = synth_audio(1, frames)[0]

    def push(obj):
        b = msgpack.packb(obj, use_single_float=True)
        buf = (C.c_uint8 * len(b)).from_buffer_copy(b)
        check(lib.dsm_session_push_msg(sess, buf, len(b)))

    push({"type": "Init"})
    push({"type": "Marker", "id": 3})
    # arbitrary packetisation: 700-sample messages, then one long one
    cut = 10 * 1920
    for o in range(0, cut, 700):
        push({"type": "Audio", "pcm": audio[o:min(o + 700, cut)].tolist()})
    push({"type": "Audio", "pcm": audio[cut:].tolist()})
    assert lib.dsm_session_steps(sess) == frames
    got = []
    buf = (C.c_uint8 * 8192)()
    n, ok = C.c_size_t(0), C.c_int32(0)
    while True:
        check(lib.dsm_session_poll_msg(sess, buf, 8192, C.byref(n), C.byref(ok)))
        if not ok.value:
            break
        got.append(msgpack.unpackb(bytes(buf[:n.value])))
    # the same frames by hand
    want, toks, codes = [], [], []
    probe = api.Mimi(mspec, mw, 1, decoder=False)
    for f in range(frames):
        fr = audio[f * 1920:(f + 1) * 1920].reshape(1, 1920)
        codes.append(probe.encode_step(fr)[0].astype(np.int64))
        for m in asr_m.step_pcm(fr):
            if m[0] == "step":
                want.append({"type": "Step", "step_idx": m[1], "prs": [float(x) for x in m[2][:, 0]], "buffered_pcm": 0})
            elif m[0] == "word":
                ids = (C.c_uint32 * len(m[1]))(*m[1])
                tb = C.create_string_buffer(1024)
                tn = C.c_size_t(0)
                check(lib.dsm_spm_decode(spm, ids, len(m[1]), tb, 1024, C.byref(tn)))
                want.append({"type": "Word", "text": tb.raw[:tn.value].decode(), "start_time": m[2]})
            else:
                want.append({"type": "EndWord", "stop_time": m[1]})
        toks.append(int(asr_m.last_tokens[0]))
    assert [g["type"] for g in got] == [w["type"] for w in want]
    for g, w in zip(got, want):
        if g["type"] == "Step":
            assert g["step_idx"] == w["step_idx"] and np.allclose(g["prs"], w["prs"], atol=1e-6)
        else:
            assert g == w
    # token log (asr.rs:126-176)
    path = tmp_path / "log.safetensors"
    check(lib.dsm_session_save_log(sess, str(path).encode()))
    raw = path.read_bytes()
    hlen = struct.unpack("<Q", raw[:8])[0]
    hdr = json.loads(raw[8:8 + hlen])
    assert hdr["text"] == {"dtype": "I64", "shape": [1, frames], "data_offsets": hdr["text"]["data_offsets"]}
    assert hdr["audio"]["shape"] == [1, mspec.n_q, frames] and hdr["audio"]["dtype"] == "I64"
    data = raw[8 + hlen:]
    text = np.frombuffer(data[slice(*hdr["text"]["data_offsets"])], np.int64)
    aud = np.frombuffer(data[slice(*hdr["audio"]["data_offsets"])], np.int64).reshape(mspec.n_q, frames)
    assert text.tolist() == toks
    assert np.array_equal(aud, np.stack(codes, axis=1))
    lib.dsm_session_destroy(sess)
    lib.dsm_spm_close(spm)


def test_tts_audio_sink_matches_oracle_decode():
    """The TTS server's consumer of decode_step (moshi-server/src/tts.rs:512-560, 917-929): delay gate, Pcm and
    PcmMessagePack framing, chunked decode of a finished token matrix -- against the oracle's decode_step."""
    import torch
    from dsm_b200 import _lib, api
    from dsm_b200._lib import check, ptr
    from oracle.mimi import Mimi as OMimi
    lib = _lib.load()
    mspec = tiny_mimi()
    mw = mimi_weights(mspec, seed=0, layer_scale=0.5, decoder=True)
    steps, delay_t, delay_a = 9, 3, 2
    rng = np.random.default_rng(2)
    codes = rng.integers(0, mspec.bins, (mspec.n_q, steps)).astype(np.uint32)

    def sink(fmt):
        m = api.Mimi(mspec, mw, 1, encoder=False, decoder=True)
        h = C.c_void_p()
        check(lib.dsm_tts_sink_create(m._h, delay_t, delay_a, fmt, C.byref(h)))
        return m, h

    o = OMimi(mspec, mw, 1, decoder=True)
    want = []
    for t in range(delay_t + delay_a, steps):
        pcm = o.decode_step(torch.from_numpy(codes[:, t].astype(np.int64)).view(1, -1, 1), None)
        want.append(pcm.reshape(-1).numpy())
    for fmt in (0, 1):
        m, h = sink(fmt)
        buf = (C.c_uint8 * 16384)()
        n = C.c_size_t(0)
        got = []
        for t in range(steps):
            frame = np.ascontiguousarray(codes[:, t])
            check(lib.dsm_tts_sink_push(h, ptr(frame) if t != 1 else None, t, buf, 16384, C.byref(n)))
            if t < delay_t + delay_a:
                assert n.value == 0
                continue
            raw = bytes(buf[:n.value])
            if fmt == 0:
                assert len(raw) == 1920 * 4
                got.append(np.frombuffer(raw, "<f4"))
            else:
                msg = msgpack.unpackb(raw)
                assert msg["type"] == "Audio" and len(msg["pcm"]) == 1920
                got.append(np.asarray(msg["pcm"], np.float32))
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert np.abs(g - w).max() / (np.abs(w).max() + 1e-9) < 1e-4
        lib.dsm_tts_sink_destroy(h)
    # the chunked decode of the whole matrix from a fresh state (tts.rs:917-929)
    m, h = sink(0)
    pcm_all = np.zeros(steps * 1920, np.float32)
    check(lib.dsm_tts_sink_decode_all(h, ptr(np.ascontiguousarray(codes)), steps, ptr(pcm_all)))
    o2 = OMimi(mspec, mw, 1, decoder=True)
    ref = np.concatenate([o2.decode_step(torch.from_numpy(codes[:, t].astype(np.int64)).view(1, -1, 1), None).reshape(-1).numpy()
                          for t in range(steps)])
    assert np.abs(pcm_all - ref).max() / np.abs(ref).max() < 1e-4
    lib.dsm_tts_sink_destroy(h)
