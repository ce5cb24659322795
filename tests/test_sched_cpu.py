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
