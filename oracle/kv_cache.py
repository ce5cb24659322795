"""Oracle: per-stream ring KV cache bookkeeping.  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Restates /root/reference/server/rust/moshi/moshi-core/src/kv_cache.rs:54-237 (ScatteredCacheBuilder)
and :20-51 (ScatteredKvCache).  Integer state is kept in int64 tensors; values are u32 in the
reference.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

NEG_INF = float("-inf")


class ScatteredCacheBuilder:
    def __init__(self, batch_size: int, context: int):
        self.batch_size, self.context = batch_size, context
        self.positions = torch.zeros(batch_size, dtype=torch.int64)   # kv_cache.rs:70
        self.indices = torch.zeros(batch_size, dtype=torch.int64)     # kv_cache.rs:71

    def reset(self):                                                 # kv_cache.rs:100-105
        self.positions.zero_()
        self.indices.zero_()

    def reset_batch_index(self, b: int):                             # kv_cache.rs:111-117
        self.positions[b] = 0
        self.indices[b] = 0

    def indices_and_mask(self, seq_len: int, batch_mask: Sequence[bool]) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns (indices ``[B, t]`` int64, mask ``[B, 1, t, context]`` float32 of 0 / -inf)."""
        if self.context <= seq_len:
            raise NotImplementedError("indices_and_mask_abs is not on the streaming path")
        if seq_len == 1:
            return self._fast(batch_mask)
        return self._slow(seq_len, batch_mask)

    def _fast(self, batch_mask):
        # kv_cache.rs:130-170
        active = torch.tensor([bool(b) for b in batch_mask])
        prev_indices = self.indices.clone()
        prev_positions = self.positions.clone()
        next_indices = self.indices + 1
        next_indices = next_indices - (next_indices >= self.context).to(torch.int64) * self.context
        next_positions = self.positions + 1
        self.indices = torch.where(active, next_indices, self.indices)
        self.positions = torch.where(active, next_positions, self.positions)
        indices = torch.where(active, prev_indices, self.indices)
        start_pos = torch.where(active, prev_positions, self.positions)
        start_pos = torch.where(~active, torch.full_like(start_pos, 0xFFFFFFFF), start_pos)
        arange = torch.arange(self.context).unsqueeze(0)
        mask_bool = arange > start_pos.unsqueeze(1)
        mask = torch.where(mask_bool, torch.tensor(NEG_INF), torch.tensor(0.0))
        return indices.unsqueeze(1), mask.unsqueeze(1).unsqueeze(1)

    def _slow(self, seq_len, batch_mask):
        # kv_cache.rs:176-237
        b, context = self.batch_size, self.context
        indices_cpu = self.indices.tolist()
        positions_cpu = self.positions.tolist()
        masks: List[float] = []
        cache_indices: List[int] = []
        next_indices = list(indices_cpu)
        next_positions = list(positions_cpu)
        big = 1 << 62
        for bi, active in enumerate(batch_mask):
            if not active:
                masks.extend([0.0] * (seq_len * context))
                cache_indices.extend([indices_cpu[bi]] * seq_len)
                continue
            start_index = next_indices[bi]
            start_pos = next_positions[bi]
            all_pos = [big] * context
            if start_pos < context:
                for i in range(start_pos):
                    all_pos[i] = i
            else:
                offset = start_pos - start_index
                for i in range(context):
                    all_pos[i] = i + offset if i < start_index else i + offset - context
            for s in range(seq_len):
                index = next_indices[bi]
                all_pos[index] = s + start_pos
                cache_indices.append(index)
                next_indices[bi] += 1
                next_positions[bi] += 1
                if next_indices[bi] >= context:
                    next_indices[bi] = 0
            for s in range(seq_len):
                my_pos = s + start_pos
                masks.extend(0.0 if p <= my_pos else NEG_INF for p in all_pos)
        self.indices = torch.tensor(next_indices, dtype=torch.int64)
        self.positions = torch.tensor(next_positions, dtype=torch.int64)
        mask = torch.tensor(masks, dtype=torch.float32).view(b, 1, seq_len, context)
        indices = torch.tensor(cache_indices, dtype=torch.int64).view(b, seq_len)
        return indices, mask


class ScatteredKvCache:
    """kv_cache.rs:20-51: k, v rings ``[B, H, context, D]`` written with scatter_set on dim 2."""

    def __init__(self, batch_size: int, num_heads: int, context: int, head_dim: int):
        self.k = torch.zeros(batch_size, num_heads, context, head_dim)
        self.v = torch.zeros(batch_size, num_heads, context, head_dim)
        self.context = context

    def append(self, k: torch.Tensor, v: torch.Tensor, indices: torch.Tensor):
        # kv_cache.rs:28-42; k, v are [B, H, t, D]; indices [B, t]
        idx = indices[:, None, :, None].expand_as(k)
        self.k.scatter_(2, idx, k)
        self.v.scatter_(2, idx, v)
        return self.k, self.v
