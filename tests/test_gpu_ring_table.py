"""The reference's own known-answer table for the ring index / mask bookkeeping
(server/rust/moshi/moshi-core/src/kv_cache.rs:338-405, test_scattered_kv_cache) run against the DEVICE ring logic:
`ring_advance_kernel` plus the visibility functions the attention kernels share (`dsm_debug_ring_trace`), and a
randomised differential against the oracle's ScatteredCacheBuilder over t = 1 / 2 / 3, inactive streams and wraps."""
import math

import numpy as np
import pytest

from test_oracle_kv_cache import REFERENCE_TABLE

pytestmark = pytest.mark.gpu


def _trace(batch, context, steps):
    """steps: list of (t, mask list) -> list of (indices [B,t], mask [B,t,ctx]) from the device."""
    from dsm_b200 import _lib
    from dsm_b200._lib import check, ptr
    lib = _lib.load()
    t = np.array([s[0] for s in steps], np.int32)
    masks = np.array([[1 if x else 0 for x in s[1]] for s in steps], np.uint8)
    n_idx = int(batch * t.sum())
    idx = np.zeros(n_idx, np.uint32)
    am = np.zeros(n_idx * context, np.float32)
    check(lib.dsm_debug_ring_trace(batch, context, len(steps), ptr(t), ptr(masks), ptr(idx), ptr(am)))
    out, o = [], 0
    for ti in t:
        n = batch * int(ti)
        out.append((idx[o:o + n].reshape(batch, ti), am[o * context:(o + n) * context].reshape(batch, ti, context)))
        o += n
    return out


def test_reference_table_on_device():
    got = _trace(2, 5, [(t, bm) for t, bm, _, _ in REFERENCE_TABLE])
    for (t, bm, exp_idx, exp_mask), (idx, mask) in zip(REFERENCE_TABLE, got):
        assert idx.tolist() == exp_idx, (t, bm)
        assert mask.tolist() == exp_mask, (t, bm)


@pytest.mark.parametrize("context,batch,seed", [(5, 3, 0), (6, 4, 1), (12, 5, 2), (250, 3, 3)])
def test_random_schedule_matches_oracle_builder(context, batch, seed):
    from oracle.kv_cache import ScatteredCacheBuilder
    rng = np.random.default_rng(seed)
    n_steps = 60 if context <= 12 else 300
    steps = []
    for _ in range(n_steps):
        t = int(rng.integers(1, min(4, context)))
        steps.append((t, [bool(x) for x in rng.random(batch) > 0.3]))
    got = _trace(batch, context, steps)
    oracle = ScatteredCacheBuilder(batch, context)
    for (t, bm), (idx, mask) in zip(steps, got):
        o_idx, o_mask = oracle.indices_and_mask(t, bm)
        assert idx.tolist() == o_idx.tolist()
        assert np.array_equal(mask, o_mask[:, 0].numpy())
