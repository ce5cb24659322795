"""Host logic of the batched step scheduler (dsm_sched_*) in dry mode (no device step): slot allocator,
per-slot PCM queues and frame gather, mask, markers, slot re-use.  Mirrors the behaviour of
moshi-server/src/batched_asr.rs:90-113 (Channel::extend_data), :527-653 (pre_process_pipelined),
:703-715 (markers), :795-808 (slot allocation)."""
import numpy as np
import pytest

from helpers import ROOT  # noqa: F401  (sys.path)
from dsm_b200 import api
from dsm_b200._lib import DsmError

FRAME = 1920


def test_slots_are_handed_out_and_refused_at_capacity():
    s = api.BatchedAsr(None, 3)
    got = [s.open() for _ in range(3)]
    assert sorted(x[0] for x in got) == [0, 1, 2]
    assert len({x[1] for x in got}) == 3                  # distinct channel ids
    assert s.free_slots() == 0
    with pytest.raises(DsmError) as e:
        s.open()
    assert e.value.code == -7                             # capacity (close code 4000 in the reference)
    s.close(1)
    s.tick()                                              # closed slots are reclaimed by the next tick
    assert s.free_slots() == 1
    slot, cid = s.open()
    assert slot == 1 and cid not in {x[1] for x in got}   # same slot, new connection


def test_frames_are_gathered_one_per_tick_with_mask():
    s = api.BatchedAsr(None, 4)
    a, _ = s.open()
    b, _ = s.open()
    rng = np.random.default_rng(0)
    xa = rng.standard_normal(FRAME * 2 + 500).astype(np.float32)
    xb = rng.standard_normal(FRAME - 1).astype(np.float32)
    s.push_audio(a, xa[:1000])
    s.push_audio(a, xa[1000:])
    s.push_audio(b, xb)
    stepped, outs = s.tick()
    assert stepped and s.last_mask().tolist() == [1, 0, 0, 0]      # b has no full frame yet
    assert {o["kind"] for o in outs} == {"ready"}                   # both new slots were reset first
    np.testing.assert_array_equal(s.batch_pcm()[a], xa[:FRAME])
    assert not s.batch_pcm()[b].any()                               # inactive rows are zero-filled
    s.push_audio(b, np.ones(1, np.float32))                        # completes b's first frame
    stepped, _ = s.tick()
    assert stepped and s.last_mask().tolist() == [1, 1, 0, 0]
    np.testing.assert_array_equal(s.batch_pcm()[a], xa[FRAME:2 * FRAME])
    np.testing.assert_array_equal(s.batch_pcm()[b][:-1], xb)
    stepped, _ = s.tick()                                           # 500 samples left in a: no frame
    assert not stepped and s.step_idx() == 2


def test_marker_fires_after_buffered_frames_plus_delay():
    s = api.BatchedAsr(None, 2)                                      # dry mode: delay 0
    a, cid = s.open()
    s.push_audio(a, np.zeros(FRAME * 3, np.float32))
    s.push_marker(a, 42)                                            # 3 frames buffered -> fires at step 3
    fired = []
    for _ in range(4):
        _, outs = s.tick()
        fired.append([o["marker_id"] for o in outs if o["kind"] == "marker"])
    assert fired == [[], [], [42], []]


def test_marker_of_a_closed_connection_is_dropped():
    s = api.BatchedAsr(None, 1)
    a, _ = s.open()
    s.push_audio(a, np.zeros(FRAME * 2, np.float32))
    s.push_marker(a, 7)
    s.tick()
    s.close(a)
    s.tick()
    a2, _ = s.open()
    assert a2 == a
    s.push_audio(a2, np.zeros(FRAME * 3, np.float32))
    seen = []
    for _ in range(3):
        seen += [o for o in s.tick()[1] if o["kind"] == "marker"]
    assert seen == []


def test_errors():
    s = api.BatchedAsr(None, 1)
    with pytest.raises(DsmError):
        s.push_audio(0, np.zeros(4, np.float32))                    # slot not open
    with pytest.raises(DsmError):
        s.push_audio(5, np.zeros(4, np.float32))                    # out of range


def test_stale_connection_is_refused():
    """A client of a closed connection must not reach whoever re-opened its slot (batched_asr.rs:110-118)."""
    s = api.BatchedAsr(None, 1)
    a, cid = s.open()
    s.push_audio(a, np.zeros(FRAME, np.float32), channel_id=cid)
    s.close(a, channel_id=cid)
    s.tick()
    a2, cid2 = s.open()
    assert a2 == a and cid2 != cid
    for call in (lambda: s.push_audio(a, np.zeros(4, np.float32), channel_id=cid),
                 lambda: s.push_marker(a, 1, channel_id=cid), lambda: s.close(a, channel_id=cid)):
        with pytest.raises(DsmError):
            call()
    s.push_audio(a2, np.zeros(4, np.float32), channel_id=cid2)


def test_marker_position_is_taken_when_it_arrives():
    """batched_asr.rs:577-585 evaluates `current_data` when the Marker message is processed in the connection's message
    order: audio pushed AFTER the marker must not delay it."""
    s = api.BatchedAsr(None, 1)
    a, _ = s.open()
    s.push_audio(a, np.zeros(FRAME * 2, np.float32))
    s.push_marker(a, 9)                                             # 2 frames ahead of it
    s.push_audio(a, np.zeros(FRAME * 5, np.float32))                # later audio
    fired = []
    for _ in range(4):
        fired.append([o["marker_id"] for o in s.tick()[1] if o["kind"] == "marker"])
    assert fired == [[], [9], [], []]


def test_queue_capacity_is_reported():
    s = api.BatchedAsr(None, 1)
    a, _ = s.open()
    s.push_audio(a, np.zeros(1 << 17, np.float32))                  # exactly the ring
    with pytest.raises(DsmError) as e:
        s.push_audio(a, np.zeros(1, np.float32))
    assert e.value.code == -7
    s.tick()
    s.push_audio(a, np.zeros(FRAME, np.float32))                    # a frame was consumed: room again


def test_multi_device_routing_and_independent_ticks():
    """Device g owns slots [g*B, (g+1)*B); connections are spread over the devices; each device steps on its own."""
    nd, B = 4, 3
    s = api.BatchedAsr(None, B, n_devices=nd)
    assert s.n_devices == nd and s.slots_per_device == B and s.free_slots() == nd * B
    opened = [s.open() for _ in range(nd * B)]
    slots = sorted(x[0] for x in opened)
    assert slots == list(range(nd * B))
    # least-loaded placement: the first nd connections land on nd different devices
    assert sorted(x[0] // B for x in opened[:nd]) == list(range(nd))
    with pytest.raises(DsmError):
        s.open()
    # only device 2's slots get audio: only device 2 steps
    for sl in range(2 * B, 3 * B):
        s.push_audio(sl, np.full(FRAME, float(sl), np.float32))
    stepped, outs = s.tick()
    assert stepped and [s.device_step_idx(g) for g in range(nd)] == [0, 0, 1, 0]
    assert sorted(o["slot"] for o in outs if o["kind"] == "ready") == list(range(nd * B))
    st = s.stats(2)
    assert st["steps"] == 1 and st["frames"] == B and s.stats(0)["frames"] == 0


def test_threaded_mode_delivers_every_frame_and_marker():
    """One tick thread per device, a pusher thread per group of connections, events drained with poll: nothing is
    lost, per-connection order is kept, the marker arrives after the frames queued ahead of it."""
    import threading
    import time
    nd, B, frames = 2, 8, 40
    s = api.BatchedAsr(None, B, n_devices=nd)
    conns = [s.open() for _ in range(nd * B)]
    s.start()
    try:
        def pusher(group):
            for f in range(frames):
                for slot, cid in group:
                    while True:
                        try:
                            s.push_audio(slot, np.zeros(FRAME, np.float32), channel_id=cid)
                            break
                        except DsmError:
                            time.sleep(0.001)               # ring full: the client waits
                if f == 9:
                    for slot, cid in group:
                        s.push_marker(slot, 1000 + slot, channel_id=cid)
        threads = [threading.Thread(target=pusher, args=(conns[i::4],)) for i in range(4)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        deadline = time.time() + 20
        while time.time() < deadline and min(s.stats(g)["frames"] for g in range(nd)) < frames * B:
            time.sleep(0.01)
        time.sleep(0.05)
    finally:
        s.stop()
    assert [s.stats(g)["frames"] for g in range(nd)] == [frames * B] * nd
    for slot, cid in conns:
        evs = []
        while True:
            e = s.poll(slot)
            if e is None:
                break
            evs.append(e)
        assert evs[0]["kind"] == "ready" and all(e["channel_id"] == cid for e in evs)
        marks = [e for e in evs if e["kind"] == "marker"]
        assert [m["marker_id"] for m in marks] == [1000 + slot]
        assert marks[0]["step_idx"] >= 10                    # 10 frames were queued ahead of it


def test_wire_messages_through_the_scheduler():
    """A connection's MessagePack traffic (batched_asr.rs:894-985): InMsg bytes in, OutMsg bytes out."""
    import ctypes as C
    import time

    import msgpack
    from dsm_b200 import _lib
    from dsm_b200._lib import check
    lib = _lib.load()
    s = api.BatchedAsr(None, 2)
    slot, cid = s.open()

    def push(obj):
        b = msgpack.packb(obj, use_single_float=True)
        buf = (C.c_uint8 * len(b)).from_buffer_copy(b)
        return lib.dsm_sched_push_msg(s._h, slot, cid, buf, len(b))

    def poll_all():
        out = []
        buf = (C.c_uint8 * 4096)()
        n, got = C.c_size_t(0), C.c_int32(0)
        while True:
            check(lib.dsm_sched_poll_msg(s._h, slot, None, buf, 4096, C.byref(n), C.byref(got)))
            if not got.value:
                return out
            out.append(msgpack.unpackb(bytes(buf[:n.value])))

    s.start()
    try:
        assert push({"type": "Audio", "pcm": [0.0] * (FRAME * 3)}) == 0
        assert push({"type": "Marker", "id": 5}) == 0
        assert push({"type": "Ping"}) == 0
        assert push({"type": "OggOpus", "data": [1, 2, 3]}) == -5          # no Opus decoder in this library
        assert push({"type": "Nope"}) == -1
        deadline = time.time() + 10
        msgs = []
        while time.time() < deadline and not any(m["type"] == "Marker" for m in msgs):
            msgs += poll_all()
            time.sleep(0.005)
        assert push({"type": "Init"}) == 0                                   # a client Init: reset + Ready again
        deadline = time.time() + 10
        while time.time() < deadline and sum(m["type"] == "Ready" for m in msgs) < 2:
            msgs += poll_all()
            time.sleep(0.005)
    finally:
        s.stop()
    assert msgs[0] == {"type": "Ready"}
    assert {"type": "Marker", "id": 5} in msgs
    assert sum(m["type"] == "Ready" for m in msgs) == 2
