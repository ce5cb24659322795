"""Known-answer table for the ring index / mask bookkeeping, restated verbatim from the reference's
own unit test: server/rust/moshi/moshi-core/src/kv_cache.rs:338-405 (test_scattered_kv_cache)."""
import math

from oracle.kv_cache import ScatteredCacheBuilder

NI = -math.inf

# (seq_len, batch_mask, expected indices, expected mask[b][t][ctx])
REFERENCE_TABLE = [
    (1, [True, False], [[0], [0]], [[[0.0, NI, NI, NI, NI]], [[0.0] * 5]]),
    (1, [True, False], [[1], [0]], [[[0.0, 0.0, NI, NI, NI]], [[0.0] * 5]]),
    (3, [False, True], [[2, 2, 2], [0, 1, 2]],
     [[[0.0] * 5, [0.0] * 5, [0.0] * 5],
      [[0.0, NI, NI, NI, NI], [0.0, 0.0, NI, NI, NI], [0.0, 0.0, 0.0, NI, NI]]]),
    (3, [True, True], [[2, 3, 4], [3, 4, 0]],
     [[[0.0, 0.0, 0.0, NI, NI], [0.0, 0.0, 0.0, 0.0, NI], [0.0] * 5],
      [[NI, 0.0, 0.0, 0.0, NI], [NI, 0.0, 0.0, 0.0, 0.0], [0.0] * 5]]),
    (1, [True, False], [[0], [1]], [[[0.0] * 5], [[0.0] * 5]]),
    (2, [True, False], [[1, 2], [1, 1]],
     [[[0.0, 0.0, NI, 0.0, 0.0], [0.0] * 5], [[0.0] * 5, [0.0] * 5]]),
]


def test_scattered_kv_cache_reference_table():
    cache = ScatteredCacheBuilder(2, 5)
    for seq_len, bm, exp_idx, exp_mask in REFERENCE_TABLE:
        idx, mask = cache.indices_and_mask(seq_len, bm)
        assert idx.tolist() == exp_idx
        assert mask[:, 0].tolist() == exp_mask
