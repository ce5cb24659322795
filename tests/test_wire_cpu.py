"""MessagePack wire messages (csrc/wire.cu) against the reference's known-answer bytes
(client/rust/kyutai-client/src/stt/protocol.rs:81-96) and the rmp_serde `.with_struct_map()` layout of InMsg / OutMsg
(moshi-server/src/asr.rs:15-34), cross-checked with the independent `msgpack` Python package."""
import ctypes as C

import msgpack
import numpy as np
import pytest

from dsm_b200 import _lib
from dsm_b200._lib import check

# protocol.rs:83-87: OutMsg::Word { text: "hello", start_time: 1.5 }
WORD_BYTES = bytes([0x83, 0xa4]) + b"type" + bytes([0xa4]) + b"Word" + bytes([0xa4]) + b"text" + bytes([0xa5]) + b"hello" + \
    bytes([0xaa]) + b"start_time" + bytes([0xcb, 0x3f, 0xf8, 0, 0, 0, 0, 0, 0])


def _encode_out(**kw):
    lib = _lib.load()
    m = _lib.WireOutC()
    keep = []
    for k, v in kw.items():
        if k == "text":
            b = v.encode()
            buf = C.create_string_buffer(b, len(b))
            keep.append(buf)
            m.text, m.text_len = C.cast(buf, C.c_void_p), len(b)
        elif k == "prs":
            arr = (C.c_float * len(v))(*v)
            keep.append(arr)
            m.prs, m.n_prs = arr, len(v)
        else:
            setattr(m, k, v)
    out = (C.c_uint8 * 4096)()
    n = C.c_size_t(0)
    check(lib.dsm_wire_encode_out(C.byref(m), out, 4096, C.byref(n)))
    return bytes(out[:n.value])


def _decode_out(b):
    lib = _lib.load()
    m = _lib.WireOutC()
    prs = (C.c_float * 64)()
    buf = (C.c_uint8 * len(b)).from_buffer_copy(b)
    check(lib.dsm_wire_decode_out(buf, len(b), C.byref(m), prs, 64))
    text = C.string_at(m.text, m.text_len).decode() if m.text_len else ""
    return m, text, [prs[i] for i in range(m.n_prs)], buf


def test_reference_known_answer_word_message():
    assert _encode_out(kind=0, text="hello", time=1.5) == WORD_BYTES
    m, text, _, _keep = _decode_out(WORD_BYTES)
    assert m.kind == 0 and text == "hello" and m.time == 1.5


def test_out_messages_match_struct_map_layout():
    cases = [
        (dict(kind=1, time=12.48), {"type": "EndWord", "stop_time": 12.48}),
        (dict(kind=2, id=-7), {"type": "Marker", "id": -7}),
        (dict(kind=2, id=70000), {"type": "Marker", "id": 70000}),
        (dict(kind=3, step_idx=1234, prs=[0.25, 0.5, 0.125, 1.0], buffered_pcm=3840),
         {"type": "Step", "step_idx": 1234, "prs": [0.25, 0.5, 0.125, 1.0], "buffered_pcm": 3840}),
        (dict(kind=4, text="no free slot é"), {"type": "Error", "message": "no free slot é"}),
        (dict(kind=5), {"type": "Ready"}),
        (dict(kind=0, text="x" * 40, time=0.0), {"type": "Word", "text": "x" * 40, "start_time": 0.0}),
    ]
    for kw, want in cases:
        b = _encode_out(**kw)
        got = msgpack.unpackb(b, raw=False, strict_map_key=False)
        assert got == want and list(got) == list(want), (kw, got)          # same keys in the same order
        # rmp_serde writes f32 for Vec<f32> and f64 for f64; integers in their smallest encoding
        if kw["kind"] == 3:
            assert b.count(b"\xca") == 4 and b"\xcd\x04\xd2" in b and b"\xcd\x0f\x00" in b
        m, text, prs, _keep = _decode_out(b)
        assert m.kind == kw["kind"]
        if "text" in kw:
            assert text == kw["text"]
        if "prs" in kw:
            assert prs == kw["prs"] and m.step_idx == 1234 and m.buffered_pcm == 3840
        if kw["kind"] == 2:
            assert m.id == kw["id"]


def _decode_in(b):
    lib = _lib.load()
    m = _lib.WireInC()
    pcm = (C.c_float * 4096)()
    data = (C.c_uint8 * 4096)()
    buf = (C.c_uint8 * len(b)).from_buffer_copy(b)
    check(lib.dsm_wire_decode_in(buf, len(b), C.byref(m), pcm, 4096, data, 4096))
    return m, [pcm[i] for i in range(m.n)] if m.kind == 2 else None, bytes(data[:m.n]) if m.kind == 3 else None


def _encode_in(kind, id=0, pcm=None, data=None):
    lib = _lib.load()
    m = _lib.WireInC()
    m.kind, m.id = kind, id
    keep = None
    if pcm is not None:
        keep = (C.c_float * len(pcm))(*pcm)
        m.pcm, m.n = keep, len(pcm)
    if data is not None:
        keep = (C.c_uint8 * len(data))(*data)
        m.data, m.n = keep, len(data)
    out = (C.c_uint8 * 65536)()
    n = C.c_size_t(0)
    check(lib.dsm_wire_encode_in(C.byref(m), out, 65536, C.byref(n)))
    return bytes(out[:n.value])


def test_in_messages_round_trip_and_layout():
    # protocol.rs:70-79: Audio { pcm: vec![0.0, -0.25, 0.5, 1.0] } round trip
    pcm = [0.0, -0.25, 0.5, 1.0]
    b = _encode_in(2, pcm=pcm)
    assert msgpack.unpackb(b) == {"type": "Audio", "pcm": pcm}
    assert b[:12] == bytes([0x82, 0xa4]) + b"type" + bytes([0xa5]) + b"Audio" and b.count(b"\xca") >= 4
    m, got, _ = _decode_in(b)
    assert m.kind == 2 and got == pcm
    assert msgpack.unpackb(_encode_in(0)) == {"type": "Init"}
    assert msgpack.unpackb(_encode_in(4)) == {"type": "Ping"}
    assert msgpack.unpackb(_encode_in(1, id=1 << 40)) == {"type": "Marker", "id": 1 << 40}
    assert _decode_in(_encode_in(1, id=-123456))[0].id == -123456
    ogg = list(b"OggS\x00\x02\xff\x80")
    b = _encode_in(3, data=ogg)
    assert msgpack.unpackb(b) == {"type": "OggOpus", "data": ogg}          # serde writes Vec<u8> as a sequence
    assert _decode_in(b)[2] == bytes(ogg)
    # a client that sends bin (serde_bytes) or f64 samples is accepted too
    assert _decode_in(msgpack.packb({"type": "OggOpus", "data": bytes(ogg)}))[2] == bytes(ogg)
    m, got, _ = _decode_in(msgpack.packb({"type": "Audio", "pcm": [0.5, -1.0]}))
    assert got == [0.5, -1.0]
    # a full 1920-sample frame uses array16
    frame = np.linspace(-1, 1, 1920).astype(np.float32).tolist()
    b = _encode_in(2, pcm=frame)
    assert b[12:16] == bytes([0xa3]) + b"pcm" and b[16] == 0xdc and len(b) == 19 + 5 * 1920
    assert _decode_in(b)[1] == frame


@pytest.mark.parametrize("bad", [b"", b"\x81\xa4type", b"\x81\xa4type\xa5Bogus", b"\x81\xa2id\x05", b"\x90"])
def test_decode_rejects_malformed(bad):
    lib = _lib.load()
    m = _lib.WireInC()
    pcm = (C.c_float * 8)()
    data = (C.c_uint8 * 8)()
    buf = (C.c_uint8 * max(len(bad), 1)).from_buffer_copy(bad.ljust(1, b"\0"))
    assert lib.dsm_wire_decode_in(buf, len(bad), C.byref(m), pcm, 8, data, 8) != 0


def test_capacity_is_reported():
    lib = _lib.load()
    m = _lib.WireOutC()
    m.kind = 5
    n = C.c_size_t(0)
    out = (C.c_uint8 * 4)()
    assert lib.dsm_wire_encode_out(C.byref(m), out, 4, C.byref(n)) == -7 and n.value == 12
