#!/usr/bin/env python
"""Tick overhead of the slot router (dsm_sched_*), per SURVEY 8(f)1.

    python tools/sched_bench.py --dry --devices 8 --slots 256 [--pushers-per 64] [--frames 200]
        host logic only (no GPU): 8 x 256 slots, one pusher thread per 64 slots, a pipelined tick thread per device
        whose "device step" is a sleep of --dry-step-us; reports gather / fan-out time per tick.
    python tools/sched_bench.py --devices N --slots 256 [--kv bf16]
        real ASR states on N GPUs of this box: stream-frames/s through the router (pusher threads -> SPSC rings ->
        pipelined ticks -> event queues) against the same frames through bare dsm_asr_step_pcm on device 0.
Prints one JSON line.  Numbers of the dry mode are host-side only; the GPU mode is not the benchmark of record
(bench.py is) -- it shows what the router adds on top of the step.
"""
import argparse
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from dsm_b200 import api
from dsm_b200._lib import DsmError

FRAME = 1920
ap = argparse.ArgumentParser()
ap.add_argument("--dry", action="store_true")
ap.add_argument("--devices", type=int, default=1)
ap.add_argument("--slots", type=int, default=256)
ap.add_argument("--pushers-per", type=int, default=64)
ap.add_argument("--frames", type=int, default=200)
ap.add_argument("--dry-step-us", type=int, default=7000)
ap.add_argument("--kv", default="bf16")
a = ap.parse_args()

nd, B = a.devices, a.slots
keep = []
if a.dry:
    sched = api.BatchedAsr(None, B, n_devices=nd)
    sched.set_dry_step_us(a.dry_step_us)
else:
    from dsm_b200.spec import MimiSpec, stt_1b_en_fr
    asrs = []
    for g in range(nd):
        m = api.Mimi(MimiSpec(), None, B, device=g, decoder=False)
        l = api.LmModel(stt_1b_en_fr(), None, B, device=g, kv_dtype=a.kv)
        l.debug_set_position(2000)
        asrs.append(api.AsrState(m, l))
        keep += [m, l]
    sched = api.BatchedAsr(asrs)

conns = [sched.open() for _ in range(nd * B)]
rng = np.random.default_rng(0)
frame = (rng.standard_normal(FRAME) * 0.1).astype(np.float32)
groups = [conns[i:i + a.pushers_per] for i in range(0, len(conns), a.pushers_per)]


def pusher(group):
    for _ in range(a.frames):
        for slot, cid in group:
            while True:
                try:
                    sched.push_audio(slot, frame, channel_id=cid)
                    break
                except DsmError:
                    time.sleep(0.0005)


def drainer(group, counts):
    n = 0
    while not stop_drain.is_set() or n:
        n = 0
        for slot, _ in group:
            while sched.poll(slot) is not None:
                n += 1
                counts[0] += 1
        if not n:
            time.sleep(0.001)


stop_drain = threading.Event()
counts = [[0] for _ in groups]
sched.start()
t0 = time.perf_counter()
threads = [threading.Thread(target=pusher, args=(g,)) for g in groups]
drains = [threading.Thread(target=drainer, args=(g, c)) for g, c in zip(groups, counts)]
for t in threads + drains:
    t.start()
for t in threads:
    t.join()
while min(sched.stats(g)["frames"] for g in range(nd)) < a.frames * B:
    time.sleep(0.002)
dt = time.perf_counter() - t0
sched.stop()
stop_drain.set()
for t in drains:
    t.join()
st = [sched.stats(g) for g in range(nd)]
line = {"mode": "dry" if a.dry else "gpu", "devices": nd, "slots_per_device": B, "pusher_threads": len(groups),
        "frames_per_slot": a.frames, "router_stream_frames_per_s": nd * B * a.frames / dt,
        "events_delivered": sum(c[0] for c in counts),
        "per_device": [{"steps": s["steps"], "ticks": s["ticks"], "idle_sleeps": s["idle_sleeps"],
                        "gather_us_per_step": s["gather_ns"] / max(s["steps"], 1) / 1e3,
                        "fanout_us_per_step": s["fanout_ns"] / max(s["steps"], 1) / 1e3,
                        "step_wait_ms_per_step": s["step_wait_ns"] / max(s["steps"], 1) / 1e6} for s in st]}
if not a.dry:
    # the same number of frames through the bare step on device 0
    import torch
    asr0 = asrs[0]
    pcm = torch.from_numpy(np.tile(frame, (B, 1))).pin_memory()
    for _ in range(5):
        asr0.step_pcm(pcm, None, want_messages=False)
    t0 = time.perf_counter()
    for _ in range(a.frames):
        asr0.step_pcm(pcm, None, want_messages=False)
    dt0 = time.perf_counter() - t0
    line["bare_step_pcm_ms"] = dt0 / a.frames * 1e3
    line["bare_stream_frames_per_s_one_device"] = B * a.frames / dt0
    line["router_ms_per_step"] = dt / a.frames * 1e3
print(json.dumps(line))
