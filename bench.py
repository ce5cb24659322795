#!/usr/bin/env python
"""Benchmark of the batched streaming-STT hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload asr|asr-2.6b|mimi|decode] [--streams B]

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
  * --impl reference: the reference's CPU path as restated in oracle/ (PyTorch f32, all host threads);
    the Rust/candle original cannot be built here (no cargo, un-vendored crates).
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
def ncu_attention_traffic(B, lspec):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of lm_attention_kernel, from the committed
    `ncu --set full` capture (profiles/r01_attn_set_full_raw.csv: stt-1b, 256 streams, full rings).  None when
    the capture does not describe this configuration."""
    import csv
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_attn_set_full_raw.csv")
    if not os.path.exists(path) or (B, lspec.num_heads, lspec.head_dim, lspec.context) != (256, 16, 128, 750):
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        rows = list(csv.reader(open(path)))
        ix = {n: i for i, n in enumerate(rows[0])}
        tot, n = 0.0, 0
        for r in rows[2:]:
            if "lm_attention_kernel" not in r[ix["Kernel Name"]] or int(r[ix["launch__grid_size"]]) != B * lspec.num_heads:
                continue
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[ix[k]]) * scale[rows[1][ix[k]]]
            n += 1
        return tot / n if n else None
    except Exception:
        return None


def oracle_throughput(workload: str, streams: int, steps: int, warmup: int):
    """Times oracle steps (PyTorch CPU f32 restatement of the reference's candle CPU path) on a bounded
    sample: `streams` streams of the same architecture.  Returns (frames_per_s, ms_per_step, cores)."""
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
    from dsm_b200.weights import mimi_weights
    mimi = Mimi(mspec, mimi_weights(mspec, seed=0, decoder=workload == "decode"), streams, decoder=workload == "decode")
    audio = synth_audio(streams, warmup + steps)
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
    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(warmup, warmup + steps):
        step(i)
    dt = time.perf_counter() - t0
    return streams * steps / dt, dt / steps * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    streams = args.ref_streams
    steps = max(1, min(args.steps, 8))
    warm = max(1, min(args.warmup, 2))
    fps, ms, cores = oracle_throughput(args.workload, streams, steps, warm)
    sample = (f"{streams} streams x {steps} steps of the {args.workload} workload (the GPU arm runs "
              f"{args.streams} streams/GPU); oracle = PyTorch CPU f32 restatement of the reference's candle CPU path")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic 24 kHz audio, random-init weights",
            "config": workload_config(args),
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
            "l2": "per-step working set (KV rings: 25 GB LM + 2 GB codec at 256 streams) far exceeds the 126 MB L2",
            "ring_prefill_steps": args.prefill}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from dsm_b200 import api
    from dsm_b200.spec import MimiSpec, stt_1b_en_fr
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

    from dsm_b200.spec import stt_2_6b_en
    B = args.streams
    mspec = MimiSpec()
    lspec = stt_2_6b_en() if args.workload == "asr-2.6b" else stt_1b_en_fr()
    is_asr = args.workload.startswith("asr")
    mimi = api.Mimi(mspec, None, B, device=local, encoder=args.workload != "decode", decoder=args.workload == "decode")
    if is_asr:
        lm = api.LmModel(lspec, None, B, device=local)
        asr = api.AsrState(mimi, lm)
        stream_ptr = lm.stream()
    else:
        lm = asr = None
        stream_ptr = api._lib.load().dsm_mimi_stream(mimi._h)
    ext = torch.cuda.ExternalStream(stream_ptr, device=local)

    # synthetic audio: a few distinct frames per stream, cycled (content does not change the work done)
    n_audio = 8
    audio = synth_audio(B, n_audio, first_stream=stream_range(rank, B)[0])
    dev_frames = [torch.from_numpy(frame(audio, i)).cuda() for i in range(n_audio)]
    host_frames = [torch.from_numpy(frame(audio, i)).pin_memory() for i in range(n_audio)]
    codes_dev = torch.empty((B, mspec.n_q), dtype=torch.int32, device="cuda")
    if args.workload == "decode":
        # config 5 input: codes ~ U[0, 2048), seed 7; device- and host-resident copies
        rng = np.random.default_rng(7 + rank)
        code_sets = [rng.integers(0, 2048, (B, mspec.n_q)).astype(np.int32) for _ in range(n_audio)]
        dev_frames = [torch.from_numpy(c).cuda() for c in code_sets]
        host_frames = [torch.from_numpy(c).pin_memory() for c in code_sets]
        pcm_dev = torch.empty((B, FRAME), dtype=torch.float32, device="cuda")
        pcm_host = torch.empty((B, FRAME), dtype=torch.float32).pin_memory()

    def step(i, src):
        if asr is not None:
            asr.step_pcm(src[i % n_audio], None, want_messages=False)
        elif args.workload == "decode":
            mimi.decode_step(src[i % n_audio], None, out=pcm_dev if src is dev_frames else pcm_host)
        else:
            mimi.encode_step(src[i % n_audio], None, out=codes_dev if src is dev_frames else None)

    def timed(n, src):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(ext)
        for i in range(n):
            step(i, src)
        e1.record(ext)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), world, "cuda")

    # fill and wrap every ring (LM 750 slots, codec 250) before any timing
    for i in range(args.prefill):
        step(i, dev_frames)
    for i in range(args.warmup):
        step(i, dev_frames)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    api.launch_count_reset()
    ms_dev = timed(args.steps, dev_frames)
    launches = api.launch_count()
    ms_e2e = timed(args.steps, host_frames)
    clocks = sampler.stop() if rank == 0 else None

    value = job_throughput(world, B, args.steps, ms_dev)
    e2e = job_throughput(world, B, args.steps, ms_e2e)

    # ---- roofline of the dominant kernel: single-query ring attention (LM), timed per launch ----
    roofline = None
    steps_done = args.prefill + args.warmup + 2 * args.steps
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
        ach = gflop / (ms_dev / args.steps)           # GFLOP per ms = TFLOP/s
        roofline = {"bound": "tensor", "kernel": "tc6_gemm_kernel (f32-accurate split-bf16 tcgen05 GEMMs of the step)",
                    "achieved": ach, "peak": tf / 6.0, "unit": "TFLOP/s", "frac": ach / (tf / 6.0), "traffic": None,
                    "peak_source": src_tf, "how": "f32 FLOPs of one step / CUDA-event step time (whole step, not one launch)"}
    if asr is not None and rank == 0:
        peak, peak_src = measured_peaks()
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
        # algorithmic bytes per launch: K and V rings read once (ctx slots x head_dim bf16 per (stream, head)),
        # plus q in and the attention output out
        d = lspec.d_model
        visible = min(steps_done, lspec.context)          # ring slots a query sees (full ring once wrapped)
        bytes_per_launch = B * lspec.num_heads * visible * lspec.head_dim * 2 * 2 + 2 * B * d * 2
        avg_ms = ms_sum / max(n_launch, 1)
        achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
        step_ms = ms_dev / args.steps
        roofline = {"bound": "hbm", "kernel": "lm_attention_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": ncu_attention_traffic(B, lspec), "peak_source": peak_src,
                    "bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms, "launches_timed": n_launch,
                    "launches_per_step": lspec.num_layers,
                    "share_of_step": avg_ms * lspec.num_layers / step_ms,
                    "how": "CUDA events around every attention launch over %d eager LM steps with full rings" % n_prof}
    if world > 1:
        dist.barrier()

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            fps, ms, cores = oracle_throughput(args.workload, args.ref_streams, 3, 1)
            cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{args.ref_streams} streams x 3 steps of the same workload on the host cores "
                             "(oracle: PyTorch CPU f32 restatement of the reference's candle CPU path)",
                   "ms_per_step": ms}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16 LM (f32 accumulate) + f32 codec",
                "data": "synthetic 24 kHz audio, seeded random-init weights",
                "config": workload_config(args),
                "realtime_streams": value / 12.5,
                "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": (B * 128 if args.workload == "decode" else B * FRAME * 4) + B, "d2h_bytes_per_step": (B * 4 + 4 * lspec.extra_heads_num * B) if asr else
                        (B * FRAME * 4 if args.workload == "decode" else B * 128)},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
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
    ap.add_argument("--prefill", type=int, default=760, help="untimed steps that fill and wrap the KV rings")
    ap.add_argument("--ref-streams", type=int, default=8, help="streams of the bounded CPU sample")
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
