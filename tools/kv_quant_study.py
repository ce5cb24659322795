#!/usr/bin/env python
"""CPU study behind the choice of the compressed LM KV ring format.

Runs the f32 oracle LM (stt-1b widths, `--layers` layers, random-init weights) next to copies whose
ScatteredKvCache stores K and V through a simulated storage format, on the same token streams, with the
ring full and wrapped, and reports per format: max relative logits error against the f32-KV oracle
(norm: max |logit|, the tolerance north_star states: 1e-2) and greedy-token agreement.

Formats:
  bf16        what the default device path stores (kv_cache.rs:88-94 in model dtype)
  int8        symmetric int8, one f32 scale per (stream, head, slot) = absmax / 127
  e4m3        raw float8 e4m3fn cast, no scale
  e4m3s       e4m3 with a per-(stream, head, slot) scale mapping absmax to 448
  int8g32     int8 with one scale per 32 elements
  int8k/int8v only K / only V in int8 (the other f32): which side the error comes from

This is a measurement aid: it imports oracle/ and is never imported by the product.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from dsm_b200.spec import stt_1b_en_fr
from dsm_b200.weights import lm_weights, round_bf16
from oracle import kv_cache as okv
from oracle.lm import LmModel


def q_bf16(x):
    return x.to(torch.bfloat16).float()


def q_int8(x):
    amax = x.abs().amax(dim=-1, keepdim=True)
    scale = amax / 127.0
    q = torch.where(scale > 0, torch.round(x / scale.clamp_min(1e-30)), torch.zeros_like(x)).clamp(-127, 127)
    return q * scale


def q_e4m3(x):
    return x.to(torch.float8_e4m3fn).float()


def q_e4m3s(x):
    amax = x.abs().amax(dim=-1, keepdim=True)
    scale = amax / 448.0
    y = torch.where(scale > 0, x / scale.clamp_min(1e-30), torch.zeros_like(x))
    return y.to(torch.float8_e4m3fn).float() * scale


def q_int8_g32(x):
    sh = x.shape
    return q_int8(x.reshape(*sh[:-1], sh[-1] // 32, 32)).reshape(sh)


def ident(x):
    return x


# name -> (K storage, V storage)
FORMATS = {"bf16": (q_bf16, q_bf16), "int8": (q_int8, q_int8), "e4m3": (q_e4m3, q_e4m3), "e4m3s": (q_e4m3s, q_e4m3s),
           "int8g32": (q_int8_g32, q_int8_g32), "int8k": (q_int8, ident), "int8v": (ident, q_int8),
           "int8k+bf16v": (q_int8, q_bf16), "bf16k+int8v": (q_bf16, q_int8)}


class QuantCache(okv.ScatteredKvCache):
    def __init__(self, base: okv.ScatteredKvCache, fn):
        self.k, self.v, self.context, self.fn = torch.zeros_like(base.k), torch.zeros_like(base.v), base.context, fn

    def append(self, k, v, indices):
        return super().append(self.fn[0](k), self.fn[1](v), indices)


def build(spec, w, B, fmt):
    lm = LmModel(spec, w, B)
    if fmt != "f32":
        for layer in lm.transformer.layers:
            layer.cache = QuantCache(layer.cache, FORMATS[fmt])
    return lm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=16)
    ap.add_argument("--context", type=int, default=48, help="ring length (wraps within --steps)")
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--streams", type=int, default=4)
    ap.add_argument("--formats", default="bf16,int8,e4m3,e4m3s")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    spec = stt_1b_en_fr().replace(num_layers=args.layers, context=args.context)
    w = {k: (torch.from_numpy(round_bf16(v)) if v.ndim == 2 else torch.from_numpy(v)) for k, v in lm_weights(spec, 0).items()}
    fmts = args.formats.split(",")
    B = args.streams
    ref = build(spec, w, B, "f32")
    models = {f: build(spec, w, B, f) for f in fmts}
    rng = np.random.default_rng(5)
    text = torch.full((B,), spec.text_start_token, dtype=torch.int64)
    worst = {f: 0.0 for f in fmts}
    agree = {f: 0 for f in fmts}
    total = 0
    t0 = time.time()
    for s in range(args.steps):
        audio = torch.from_numpy(rng.integers(0, 2048, (B, spec.audio_codebooks)))
        lg_ref, _ = ref.forward_cond(text, audio, None, [True] * B)
        tok_ref = lg_ref[:, 0].argmax(-1)
        for f in fmts:
            lg, _ = models[f].forward_cond(text, audio, None, [True] * B)
            rel = float((lg - lg_ref).abs().max() / lg_ref.abs().max())
            worst[f] = max(worst[f], rel)
            agree[f] += int((lg[:, 0].argmax(-1) == tok_ref).sum())
        total += B
        text = tok_ref          # every copy is fed the f32 oracle's token: histories stay comparable
    print(f"stt-1b widths, {args.layers} layers, ring {args.context}, {args.steps} steps x {B} streams ({time.time() - t0:.0f} s)")
    for f in fmts:
        print(f"  {f:12s}  max logits rel err {worst[f]:.3e}   greedy tokens equal {agree[f]}/{total}")


if __name__ == "__main__":
    main()
