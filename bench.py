#!/usr/bin/env python
"""Benchmark of the batched streaming-STT hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload asr|asr-2.6b|mimi|decode] [--streams B] [--kv bf16|int8|int8k]
                    [--latency-steps L] [--live-sweep B1,B2,...]

Workload (config.workload): stt-1b-en_fr full ASR step -- Mimi encode_step of the current 80 ms frame
plus the delayed-streams LM step on the previous frame's codes, greedy text token -- for 256
independent streams per GPU (BASELINE.json configs[2]).  The other single-GPU configurations are
available as --workload mimi (configs[1]: codec encode alone, 64 streams), asr-2.6b (configs[3]:
stt-2.6b-en, 128 streams) and decode (configs[4]: Mimi decode_step, 64 streams).  One "step" = one frame for every stream of the batch; metric = stream-frames/s
over all GPUs.  Synthetic 24 kHz audio (SURVEY.md 8(d) generator), seeded random-init weights of the
named architecture (no checkpoints offline).

Arms:
  * default: this repo's CUDA path through the C ABI (dsm_asr_step_pcm).  `value` is measured with the
    frame's PCM already resident in HBM, `e2e` with PCM in pinned host memory (H2D inside the timed
    region); both include the per-step device->host read of the sampled tokens.
  * --impl reference: the reference's CPU path as restated in oracle/ (PyTorch f32, all host threads) on the SAME
    number of streams; the Rust/candle original cannot be built here (no cargo, un-vendored crates).

Beyond the base contract the line carries: `roofline.step_frac` (SURVEY 8(d) algorithmic bytes of the whole step /
step time / HBM peak), `latency` (p50 / p99 of single steps through the public call with host PCM, over
--latency-steps steps: the metric's "real-time" clause is p99 <= 80 ms), `realtime_streams_live` (the batch, if it
meets that clause) and, with --live-sweep, `live_streams_max` (largest swept batch that fits and meets it).
`config.sm_partition` says whether the ASR state put its LM and codec branches on disjoint SM sets (CUDA green contexts,
DESIGN.md section 6; DSM_SM_SPLIT=0 in the environment turns that off) -- `roofline.frac` and `lm_step_ms` are then
measured inside the LM branch's partition.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

FRAME = 1920
METRIC = "asr_stream_frames_per_s"
UNIT = "frames/s"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def stream_range(rank: int, streams_per_gpu: int):
    """Streams [first, last) owned by `rank`: a disjoint slice per GPU, no step-path exchange."""
    return rank * streams_per_gpu, (rank + 1) * streams_per_gpu


def max_over_ranks(ms: float, world: int, device="cpu") -> float:
    """The job finishes when its slowest rank does: MAX-reduce of the device-timed duration."""
    if world <= 1:
        return float(ms)
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_throughput(world: int, streams_per_gpu: int, steps: int, ms_max: float) -> float:
    """Whole-job stream-frames/s: every rank processed streams_per_gpu * steps frames in ms_max."""
    return world * streams_per_gpu * steps / (ms_max * 1e-3)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])), mx.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on the host cores
# ---------------------------------------------------------------------------------------------
def _sha16(path):
    import hashlib
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()[:16]


def ncu_attention_traffic(B, lspec, kv):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the attention kernel from the committed
    `ncu --set full` capture summary (profiles/attn_traffic.json, written by tools/ncu_traffic.py).  Only returned when
    the capture describes this configuration AND the kernel source it was taken from is byte-identical to the one in
    this tree (sha of csrc/lm_kernels.cu); otherwise None -- a stale number is worse than none."""
    path = os.path.join(ROOT, "profiles", "attn_traffic.json")
    try:
        with open(path) as f:
            recs = json.load(f)
        src = _sha16(os.path.join(ROOT, "delayed-streams-modeling_b200", "csrc", "lm_kernels.cu"))
        for r in recs:
            if (r["streams"], r["heads"], r["head_dim"], r["context"], r["kv"]) == \
                    (B, lspec.num_heads, lspec.head_dim, lspec.context, kv) and r["lm_kernels_sha16"] == src:
                return float(r["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def oracle_throughput(workload: str, streams: int, steps: int, warmup: int, lm_context: int = 0, budget_s: float = 0.0):
    """Times oracle steps (PyTorch CPU f32 restatement of the reference's candle CPU path) on `streams` streams of
    the same architecture.  `lm_context` > 0 shortens the LM ring (the f32 oracle's 750-slot KV for 256 streams is
    100 GB of host memory; attention is ~6 % of the CPU step's FLOPs, so a short ring flatters the CPU arm slightly).
    `budget_s` > 0 trims `steps` so the timed loop fits.  Returns (frames_per_s, ms_per_step, cores, steps_run)."""
    import torch
    from dsm_b200.spec import MimiSpec, stt_1b_en_fr
    from dsm_b200.synth import frame, synth_audio
    from oracle.lm import AsrState, LmModel
    from oracle.mimi import Mimi

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)

    def rand_like(shapes):
        return {k: torch.randn(*s, generator=g) * (1.0 / max(s[-1], 1)) ** 0.5 for k, s in shapes.items()}

    from dsm_b200.spec import stt_2_6b_en
    mspec = MimiSpec()
    lspec = stt_2_6b_en() if workload == "asr-2.6b" else stt_1b_en_fr()
    if lm_context > 0:
        lspec = lspec.replace(context=min(lspec.context, lm_context))
    from dsm_b200.weights import mimi_weights
    mimi = Mimi(mspec, mimi_weights(mspec, seed=0, decoder=workload == "decode"), streams, decoder=workload == "decode")
    audio = synth_audio(streams, min(warmup + steps, 8))
    n_audio = audio.shape[1] // FRAME
    _frame = frame

    def frame(a, i):            # cycle a few distinct frames (content does not change the work done)
        return _frame(a, i % n_audio)
    if workload == "mimi":
        def step(i):
            mimi.encode_step(torch.from_numpy(frame(audio, i)).unsqueeze(1), None)
    elif workload == "decode":
        codes = torch.randint(0, 2048, (streams, 32, 1), generator=g)

        def step(i):
            mimi.decode_step(codes, None)
    else:
        d, h = lspec.d_model, lspec.hidden
        shapes = {"text_emb.weight": (lspec.text_in_vocab_size, d), "out_norm.alpha": (1, 1, d),
                  "text_linear.weight": (lspec.text_out_vocab_size, d)}
        for i in range(lspec.audio_codebooks):
            shapes[f"emb.{i}.weight"] = (lspec.audio_vocab_size, d)
        for l in range(lspec.num_layers):
            p = f"transformer.layers.{l}"
            shapes.update({p + ".self_attn.in_proj_weight": (3 * d, d), p + ".self_attn.out_proj.weight": (d, d),
                           p + ".norm1.alpha": (1, 1, d), p + ".norm2.alpha": (1, 1, d),
                           p + ".gating.linear_in.weight": (2 * h, d), p + ".gating.linear_out.weight": (d, h)})
        for i in range(lspec.extra_heads_num):
            shapes[f"extra_heads.{i}.weight"] = (lspec.extra_heads_dim, d)
        lm = LmModel(lspec, rand_like(shapes), streams)
        asr = AsrState(streams, lspec.asr_delay_in_tokens, mimi, lm)

        def step(i):
            asr.step_pcm(torch.from_numpy(frame(audio, i)).unsqueeze(1), None)
    tw = time.perf_counter()
    for i in range(warmup):
        step(i)
    tw = (time.perf_counter() - tw) / max(warmup, 1)
    if budget_s > 0 and warmup > 0:
        steps = max(1, min(steps, int(budget_s / max(tw, 1e-6))))
    t0 = time.perf_counter()
    for i in range(warmup, warmup + steps):
        step(i)
    dt = time.perf_counter() - t0
    return streams * steps / dt, dt / steps * 1e3, cores, steps


CPU_LM_CONTEXT = 64      # LM ring of the CPU arm (see oracle_throughput)


def run_reference(args):
    """Reference arm: the oracle on the host cores, all threads, on the SAME stream count as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    streams = args.ref_streams if args.ref_streams > 0 else args.streams
    fps, ms, cores, steps = oracle_throughput(args.workload, streams, args.steps, args.warmup, CPU_LM_CONTEXT, budget_s=150.0)
    cfg = workload_config(args)
    cfg["streams_per_gpu"] = streams
    cfg["cpu_lm_context"] = CPU_LM_CONTEXT
    cfg.pop("ring_prefill_steps", None)
    cfg.pop("lm_kv", None)
    sample = (f"{streams} streams x {steps} steps of the {args.workload} workload, LM ring {CPU_LM_CONTEXT} slots (an f32 "
              "750-slot KV is 0.39 GB per stream of host memory); oracle = PyTorch CPU f32 restatement of the "
              "reference's candle CPU path")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic 24 kHz audio, random-init weights",
            "config": cfg,
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


WORKLOADS = {
    "asr": ("stt-1b-en_fr full ASR step (Mimi encode_step + delayed-streams LM step, greedy token)", 256),
    "asr-2.6b": ("stt-2.6b-en full ASR step (Mimi encode_step + delayed-streams LM step, greedy token)", 128),
    "mimi": ("Mimi streaming encode_step + RVQ only (codec kernels in isolation)", 64),
    "decode": ("Mimi streaming decode_step (RVQ dequant + conv-transpose SEANet decoder)", 64),
}


def workload_config(args):
    return {"workload": WORKLOADS[args.workload][0], "streams_per_gpu": args.streams, "frame_samples": FRAME, "n_q": 32,
            "parallelism": f"streams sharded over {args.gpus} GPU(s), no step-path collective",
            "l2": "inputs larger than L2: the per-step working set (KV rings: 25 GB LM + 2 GB codec at 256 streams) "
                  "far exceeds the 126 MB L2, no flush needed",
            "lm_kv": args.kv, "ring_prefill_steps": args.prefill}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
MIMI_WEIGHT_BYTES = {"encode": 223.5e6, "decode": 231.9e6}
MIMI_STREAM_BYTES = {"encode": 7680 + 128 + 2 * 46104 + 8192000 + 65536,      # pcm, codes, conv state r+w, KV r, KV w
                     "decode": 128 + 7680 + 2 * 50700 + 8192000 + 65536}


def lm_streamed_weight_bytes(lspec) -> float:
    """bf16 weights every step reads once: the layers, the text head, the extra heads -- not the embedding
    tables (33 gathered rows per stream)."""
    d, h = lspec.d_model, lspec.hidden
    per_layer = 3 * d * d + d * d + 2 * h * d + d * h
    return 2.0 * (lspec.num_layers * per_layer + lspec.text_out_vocab_size * d + lspec.extra_heads_num * lspec.extra_heads_dim * d)


def step_algorithmic_bytes(workload, B, lspec, lm_kv_bytes):
    """SURVEY.md 8(d) algorithmic bytes of ONE step at B streams: (total, {part: bytes})."""
    parts = {}
    if workload.startswith("asr"):
        kv_elem = lm_kv_bytes / (B * lspec.num_layers * lspec.num_heads * lspec.context * lspec.head_dim * 2.0)
        parts["lm_weights"] = lm_streamed_weight_bytes(lspec)
        parts["lm_kv_read"] = float(lm_kv_bytes)
        parts["lm_kv_write"] = B * lspec.num_layers * 2.0 * lspec.d_model * kv_elem
        parts["lm_embeddings"] = B * (lspec.audio_codebooks + 1) * lspec.d_model * 2.0
    side = "decode" if workload == "decode" else "encode"
    parts["mimi_weights"] = MIMI_WEIGHT_BYTES[side]
    parts["mimi_streams"] = float(B) * MIMI_STREAM_BYTES[side]
    return sum(parts.values()), parts


def percentile(xs, q):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(round(q * (len(xs) - 1))))]


def run_ours(args):
    import torch
    import torch.distributed as dist

    from dsm_b200 import api
    from dsm_b200.spec import MimiSpec, stt_1b_en_fr, stt_2_6b_en
    from dsm_b200.synth import frame, synth_audio

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mspec = MimiSpec()
    lspec = stt_2_6b_en() if args.workload == "asr-2.6b" else stt_1b_en_fr()
    is_asr = args.workload.startswith("asr")
    n_audio = 8

    class Rig:
        """Handles + inputs of one batch size."""

        def __init__(self, B):
            self.B = B
            self.mimi = api.Mimi(mspec, None, B, device=local, encoder=args.workload != "decode", decoder=args.workload == "decode")
            if is_asr:
                self.lm = api.LmModel(lspec, None, B, device=local, kv_dtype=args.kv)
                self.asr = api.AsrState(self.mimi, self.lm)
                stream_ptr = self.lm.stream()
            else:
                self.lm = self.asr = None
                stream_ptr = api._lib.load().dsm_mimi_stream(self.mimi._h)
            self.ext = torch.cuda.ExternalStream(stream_ptr, device=local)
            # synthetic audio: a few distinct frames per stream, cycled (content does not change the work done)
            audio = synth_audio(B, n_audio, first_stream=stream_range(rank, B)[0])
            self.dev = [torch.from_numpy(frame(audio, i)).cuda() for i in range(n_audio)]
            self.host = [torch.from_numpy(frame(audio, i)).pin_memory() for i in range(n_audio)]
            self.codes_dev = torch.empty((B, mspec.n_q), dtype=torch.int32, device="cuda")
            if args.workload == "decode":
                # config 5 input: codes ~ U[0, 2048), seed 7; device- and host-resident copies
                rng = np.random.default_rng(7 + rank)
                code_sets = [rng.integers(0, 2048, (B, mspec.n_q)).astype(np.int32) for _ in range(n_audio)]
                self.dev = [torch.from_numpy(c).cuda() for c in code_sets]
                self.host = [torch.from_numpy(c).pin_memory() for c in code_sets]
                self.pcm_dev = torch.empty((B, FRAME), dtype=torch.float32, device="cuda")
                self.pcm_host = torch.empty((B, FRAME), dtype=torch.float32).pin_memory()

        def step(self, i, src):
            if self.asr is not None:
                self.asr.step_pcm(src[i % n_audio], None, want_messages=False)
            elif args.workload == "decode":
                self.mimi.decode_step(src[i % n_audio], None, out=self.pcm_dev if src is self.dev else self.pcm_host)
            else:
                self.mimi.encode_step(src[i % n_audio], None, out=self.codes_dev if src is self.dev else None)

        def timed(self, n, src):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record(self.ext)
            for i in range(n):
                self.step(i, src)
            e1.record(self.ext)
            barrier()
            return max_over_ranks(e0.elapsed_time(e1), world, "cuda")

        def latency(self, n):
            """Wall time of single steps through the public call with host PCM (H2D + graph + D2H of the tokens)."""
            ms = []
            for i in range(n):
                t0 = time.perf_counter()
                self.step(i, self.host)
                if self.asr is None:
                    torch.cuda.synchronize()
                ms.append((time.perf_counter() - t0) * 1e3)
            return ms

        def close(self):
            for h in (self.asr, self.lm, self.mimi):
                if h is not None:
                    h.close()

    B = args.streams
    rig = Rig(B)
    asr, lm = rig.asr, rig.lm

    # fill and wrap every ring (LM 750 slots, codec 250) before any timing
    for i in range(args.prefill):
        rig.step(i, rig.dev)
    for i in range(args.warmup):
        rig.step(i, rig.dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    api.launch_count_reset()
    ms_dev = rig.timed(args.steps, rig.dev)
    launches = api.launch_count()
    ms_e2e = rig.timed(args.steps, rig.host)
    clocks = sampler.stop() if rank == 0 else None

    value = job_throughput(world, B, args.steps, ms_dev)
    e2e = job_throughput(world, B, args.steps, ms_e2e)
    step_ms = ms_dev / args.steps

    # ---- per-step latency: the metric's "real-time" clause is p99 <= 80 ms (one 12.5 Hz frame) ----
    lat = rig.latency(args.latency_steps) if args.latency_steps > 0 else []
    latency = None
    if lat:
        p50, p99, worst = percentile(lat, 0.5), percentile(lat, 0.99), max(lat)
        p99_all = max_over_ranks(p99, world, "cuda")
        latency = {"steps": len(lat), "p50_ms": p50, "p99_ms": p99_all, "max_ms": worst, "budget_ms": 80.0,
                   "how": "host wall clock around single dsm_asr_step_pcm calls with pinned host PCM (H2D, graph, D2H of the "
                          "tokens); p99 is the max over ranks"}

    # ---- roofline ----
    peak, peak_src = measured_peaks()
    kv_bytes = lm.kv_bytes if lm is not None else 0
    step_bytes, step_parts = step_algorithmic_bytes(args.workload, B, lspec, kv_bytes)
    step_frac = step_bytes / (step_ms * 1e-3) / 1e9 / peak
    roofline = None
    steps_done = args.prefill + args.warmup + 2 * args.steps + len(lat)
    if asr is None and rank == 0:
        # codec-only workloads are tensor-pipe work: 0.468 (encode) / 0.441 (decode) GFLOP of f32 MACs per
        # stream-frame, each f32 MAC costing six bf16 MMAs in the split-bf16 tcgen05 GEMM
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                tf = float(json.load(f)["bf16_tflops_sustained"])
            src_tf = "measured (MEASURED_PEAKS.json bf16_tflops_sustained) / 6"
        except Exception:
            tf, src_tf = 1400.0, "fallback (B200_PROFILING.md 1.4 PFLOP/s sustained) / 6"
        gflop = (0.441 if args.workload == "decode" else 0.468) * B
        ach = gflop / step_ms           # GFLOP per ms = TFLOP/s
        roofline = {"bound": "tensor", "kernel": "tc6_gemm_kernel (f32-accurate split-bf16 tcgen05 GEMMs of the step)",
                    "achieved": ach, "peak": tf / 6.0, "unit": "TFLOP/s", "frac": ach / (tf / 6.0), "traffic": None,
                    "peak_source": src_tf, "how": "f32 FLOPs of one step / CUDA-event step time (whole step, not one launch)"}
    if asr is not None and rank == 0:
        api.set_graphs(False)
        lm.profile_attention(True)
        codes_host = np.random.default_rng(7).integers(0, 2048, (B, lspec.audio_codebooks)).astype(np.uint32)
        asr.step_tokens(codes_host, None, want_messages=False)      # eager warm-up
        lm.profile_read()
        n_prof = 4
        for _ in range(n_prof):
            asr.step_tokens(codes_host, None, want_messages=False)
        ms_sum, n_launch = lm.profile_read()
        lm.profile_attention(False)
        api.set_graphs(True)
        # algorithmic bytes per launch: the layer's K and V rings read once (visible slots only; for the 8-bit formats
        # including their per-slot scales), plus q in and the attention output out
        d = lspec.d_model
        visible = min(steps_done, lspec.context)          # ring slots a query sees (full ring once wrapped)
        ring_bytes = kv_bytes / lspec.num_layers * visible / lspec.context
        bytes_per_launch = ring_bytes + 2 * B * d * 2
        avg_ms = ms_sum / max(n_launch, 1)
        achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
        kname = "lm_attention_kernel" if args.kv == "bf16" else "lm_attention_q8_kernel"
        roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": ncu_attention_traffic(B, lspec, args.kv), "peak_source": peak_src,
                    "bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms, "launches_timed": n_launch,
                    "launches_per_step": lspec.num_layers,
                    "share_of_step": avg_ms * lspec.num_layers / step_ms,
                    "how": "CUDA events around every attention launch over %d eager LM steps with full rings" % n_prof}
    if asr is not None:
        # the LM decode step on its own (north_star: "the decode step >= 70 % of HBM roofline"): the same frame graph
        # without the codec branch, fed codes; full rings; device-timed over the same number of steps
        codes_dev_lm = torch.from_numpy(np.random.default_rng(7).integers(0, 2048, (B, lspec.audio_codebooks)).astype(np.int32)).cuda()
        for _ in range(3):
            asr.step_tokens(codes_dev_lm, None, want_messages=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(rig.ext)
        for _ in range(args.steps):
            asr.step_tokens(codes_dev_lm, None, want_messages=False)
        e1.record(rig.ext)
        barrier()
        lm_ms = max_over_ranks(e0.elapsed_time(e1), world, "cuda") / args.steps
        if roofline is not None:
            lm_bytes = sum(v for k, v in step_parts.items() if k.startswith("lm_"))
            roofline["lm_step_ms"] = lm_ms
            roofline["lm_step_bytes"] = lm_bytes
            roofline["lm_step_frac"] = lm_bytes / (lm_ms * 1e-3) / 1e9 / peak
    if roofline is not None:
        # the whole step against the HBM roofline (SURVEY 8(d) bytes; `frac` above is the dominant kernel alone)
        roofline["step_frac"] = step_frac
        roofline["step_bytes"] = step_bytes
        roofline["step_bytes_parts"] = step_parts
        roofline["step_achieved_gbs"] = step_bytes / (step_ms * 1e-3) / 1e9
    if world > 1:
        dist.barrier()

    sm_part = asr.sm_partition() if is_asr else (0, 0)
    # ---- optional: largest batch that fits in HBM and keeps p99 <= 80 ms ----
    live = None
    if args.live_sweep and world == 1:
        rig.close()
        del rig, asr, lm
        torch.cuda.empty_cache()
        live = {"budget_ms": 80.0, "tried": []}
        best = 0
        for b in [int(x) for x in args.live_sweep.split(",") if x]:
            rec = {"streams": b}
            try:
                r = Rig(b)
                if r.lm is not None:
                    r.lm.debug_set_position(2 * lspec.context)       # full rings without replaying 750 steps
                for i in range(10):
                    r.step(i, r.host)
                ms = r.latency(200)
                rec.update(p50_ms=percentile(ms, 0.5), p99_ms=percentile(ms, 0.99), fits=True,
                           lm_kv_gb=(r.lm.kv_bytes if r.lm else 0) / 1e9, mimi_gb=r.mimi.device_bytes / 1e9)
                if rec["p99_ms"] <= 80.0:
                    best = max(best, b)
                r.close()
                del r
            except Exception as e:          # out of memory at create: does not fit
                rec.update(fits=False, error=str(e)[:120])
            torch.cuda.empty_cache()
            live["tried"].append(rec)
        live["live_streams_max"] = best

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            ref_b = args.ref_streams if args.ref_streams > 0 else B
            fps, ms, cores, n = oracle_throughput(args.workload, ref_b, 2, 1, CPU_LM_CONTEXT)
            cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{ref_b} streams x {n} steps of the same workload on the host cores, LM ring {CPU_LM_CONTEXT} slots "
                             "(oracle: PyTorch CPU f32 restatement of the reference's candle CPU path)",
                   "ms_per_step": ms}
        rt_ok = latency is not None and latency["p99_ms"] <= 80.0
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": ("bf16 LM (f32 accumulate, %s KV ring) + f32 codec" % args.kv) if is_asr else "f32 codec (split-bf16 tensor-core GEMMs)",
                "data": "synthetic 24 kHz audio, seeded random-init weights",
                "config": workload_config(args),
                # throughput equivalent (frames/s / 12.5) and the live count: this batch, if every step met the frame budget
                "realtime_streams_equiv": value / 12.5,
                "realtime_streams_live": (world * B if rt_ok else 0) if latency is not None else None,
                "latency": latency,
                "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": (B * 128 if args.workload == "decode" else B * FRAME * 4) + B,
                        "d2h_bytes_per_step": (B * 4 + 4 * lspec.extra_heads_num * B) if is_asr else
                        (B * FRAME * 4 if args.workload == "decode" else B * 128)},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        if is_asr:
            lm_sms, codec_sms = sm_part
            line["config"]["sm_partition"] = ({"lm_sms": lm_sms, "codec_sms": codec_sms,
                                               "how": "the LM and codec branches of the frame graph run on disjoint SM sets (CUDA green contexts); "
                                                      "roofline.frac / lm_step_ms are measured inside the LM partition"}
                                              if lm_sms else None)
        if live is not None:
            line["live_streams_max"] = live["live_streams_max"]
            line["live_sweep"] = live
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="asr", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=0, help="streams per GPU (default: the workload's BASELINE batch)")
    ap.add_argument("--kv", default="bf16", choices=["bf16", "int8", "int8k"], help="LM KV ring storage (bf16 = the reference's)")
    ap.add_argument("--prefill", type=int, default=760, help="untimed steps that fill and wrap the KV rings")
    ap.add_argument("--latency-steps", type=int, default=500, help="single steps timed for p50 / p99 (0 = skip)")
    ap.add_argument("--live-sweep", default="", help="comma-separated batch sizes to try for live_streams_max (1 GPU)")
    ap.add_argument("--ref-streams", type=int, default=0, help="streams of the CPU arm (default: the same as --streams)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.streams <= 0:
        args.streams = WORKLOADS[args.workload][1]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
