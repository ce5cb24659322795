"""Oracle: delayed-streams LM step and the batched ASR state.  TEST INFRASTRUCTURE -- see
oracle/__init__.py.

Restates /root/reference/server/rust/moshi/moshi-core/src/lm.rs:797-1008 (LmModel without depformer)
and asr.rs:9-284 (asr::State), float32 like the reference's CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from dsm_b200.spec import LmSpec

from .transformer import StreamingTransformer, TransformerConfig, rms_norm


def _as_torch(w) -> Dict[str, torch.Tensor]:
    return {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))).float()
            for k, v in w.items()}


class LmModel:
    def __init__(self, spec: LmSpec, weights, batch_size: int):
        w = _as_torch(weights)
        self.spec, self.batch_size = spec, batch_size
        cfg = TransformerConfig(spec.d_model, spec.num_heads, spec.num_layers, spec.dim_feedforward,
                                spec.context, spec.max_period, norm="rms", gating=True, layer_scale=False,
                                eps=spec.rms_eps)
        self.transformer = StreamingTransformer(batch_size, cfg, w, "transformer")
        self.text_emb = w["text_emb.weight"]
        self.audio_embs = [w[f"emb.{i}.weight"] for i in range(spec.audio_codebooks)]
        self.out_norm = w["out_norm.alpha"].reshape(-1)
        self.text_linear = w["text_linear.weight"]
        self.extra = [w[f"extra_heads.{i}.weight"] for i in range(spec.extra_heads_num)]

    def forward_cond(self, text_ids: torch.Tensor, audio_ids: torch.Tensor, cond: Optional[torch.Tensor],
                     mask: Sequence[bool]) -> Tuple[torch.Tensor, torch.Tensor]:
        """lm.rs:957-1008.  text_ids ``[B]``, audio_ids ``[B, n_cb]`` -> (logits ``[B, 1, V]``,
        ys ``[B, 1, D]``)."""
        emb = self.text_emb[text_ids].unsqueeze(1)
        for i, table in enumerate(self.audio_embs):
            emb = emb + table[audio_ids[:, i]].unsqueeze(1)
        if cond is not None:
            emb = emb + cond.view(1, 1, -1)
        ys = self.transformer.forward(emb, mask)
        ys = rms_norm(ys, self.out_norm, self.spec.rms_eps)
        return F.linear(ys, self.text_linear), ys

    def extra_heads(self, ys: torch.Tensor) -> List[torch.Tensor]:
        return [F.linear(ys, h) for h in self.extra]      # lm.rs:948-955

    def reset_state(self):
        self.transformer.reset_state()

    def reset_batch_idx(self, b: int):
        self.transformer.reset_batch_idx(b)


@dataclass
class ItemState:                         # asr.rs:16-51
    text_token: int
    step_idx: int = 0
    word_tokens: List[int] = field(default_factory=list)
    unended_word: bool = False
    last_stop_time: float = 0.0


class AsrState:
    """asr.rs:53-284.  Messages are tuples: ('step', step_idx, prs), ('word', tokens, start_time,
    batch_idx), ('end_word', stop_time, batch_idx)."""

    def __init__(self, batch_size: int, asr_delay_in_tokens: int, mimi, lm: LmModel):
        self.lm, self.mimi = lm, mimi
        self.asr_delay_in_tokens = asr_delay_in_tokens
        self.model_step_idx = 0
        self.batch = [ItemState(lm.spec.text_start_token) for _ in range(batch_size)]
        self.next_codebooks = torch.full((batch_size, lm.spec.audio_codebooks), lm.spec.audio_pad_token,
                                         dtype=torch.int64)
        self.last_logits: Optional[torch.Tensor] = None
        self.last_tokens: Optional[torch.Tensor] = None
        # Checker aid (not reference behaviour): tokens [B] sampled by the implementation under test for
        # the step about to run.  Where such a token is a near-tie of the oracle's own argmax (logit gap
        # <= tie_tol * max|logit|, i.e. inside the logits tolerance) the oracle follows it, so that one
        # legitimate tie flip does not make every later frame incomparable.  `ties_followed` counts them.
        self.follow_tokens: Optional[torch.Tensor] = None
        self.tie_tol = 0.0
        self.ties_followed = 0

    def step_pcm(self, pcm: torch.Tensor, mask=None, cond=None):
        # asr.rs:115-131
        codes = self.mimi.encode_step(pcm, mask)
        if codes is None:
            return []
        return self.step_tokens(codes, mask, cond)

    def step_tokens(self, audio_tokens: torch.Tensor, mask=None, cond=None):
        # asr.rs:147-255
        b, codebooks, steps = audio_tokens.shape
        assert b == len(self.batch)
        active = [True] * b if mask is None else [bool(m) for m in mask]
        tmask = torch.tensor(active)
        msgs = []
        is_first = torch.tensor([s.step_idx == 0 for s in self.batch])
        pad = self.lm.spec.audio_pad_token
        for step in range(steps):
            cur = audio_tokens[:, :, step].to(torch.int64)
            next_tokens = torch.where(is_first.view(b, 1), torch.full_like(cur, pad), self.next_codebooks)
            self.next_codebooks = torch.where(tmask.view(b, 1), cur, self.next_codebooks)
            text = torch.tensor([self.lm.spec.text_start_token if s.step_idx == 0 else s.text_token
                                 for s in self.batch], dtype=torch.int64)
            logits, ys = self.lm.forward_cond(text, next_tokens, cond, active)
            self.model_step_idx += 1
            heads = self.lm.extra_heads(ys)
            if heads:
                pr = torch.softmax(torch.stack(heads, 0).float(), dim=-1)[:, :, 0, 0]
                msgs.append(("step", self.model_step_idx, pr.clone()))
            tokens = logits[:, 0].argmax(dim=-1)
            if self.follow_tokens is not None:
                lg = logits[:, 0].float()
                alt = torch.as_tensor(self.follow_tokens, dtype=torch.int64)
                gap = lg.gather(1, tokens.view(b, 1)) - lg.gather(1, alt.view(b, 1))
                near = (gap[:, 0] <= self.tie_tol * lg.abs().max()) & (alt != tokens) & tmask
                self.ties_followed += int(near.sum())
                tokens = torch.where(near, alt, tokens)
                self.follow_tokens = None
            self.last_logits, self.last_tokens = logits[:, 0], tokens
            for bi, tok in enumerate(tokens.tolist()):
                if not active[bi]:
                    continue
                item = self.batch[bi]
                item.text_token = tok
                item.step_idx += 1
                is_first[bi] = False
                if item.step_idx >= self.asr_delay_in_tokens:
                    if tok in (0, 3):
                        if item.word_tokens:
                            msgs.append(("word", list(item.word_tokens), item.last_stop_time, bi))
                            item.word_tokens = []
                            item.unended_word = True
                    else:
                        item.word_tokens.append(tok)
                    if tok == 0:
                        stop = (item.step_idx - self.asr_delay_in_tokens) / 12.5
                        if item.unended_word:
                            item.unended_word = False
                            msgs.append(("end_word", stop, bi))
                        item.last_stop_time = stop
        return msgs

    def reset_batch_idx(self, bi: int):
        # asr.rs:257-266
        self.batch[bi] = ItemState(self.lm.spec.text_start_token)
        self.lm.reset_batch_idx(bi)
        self.mimi.reset_batch_idx(bi)
