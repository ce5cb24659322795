"""CPU oracle for the batched streaming-STT hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in eager PyTorch float32 on the CPU, the algorithm that the reference
(``/root/reference/server/rust/moshi/moshi-core/src``) runs on its candle CPU path.  It is the checker
the CUDA path is compared against and the CPU baseline ``bench.py`` times; nothing in the product
package (``delayed-streams-modeling_b200/``) imports it, and the product fails loudly when its CUDA
library is missing rather than falling back to this code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.

Why a restatement: the reference is Rust on top of un-vendored crates (candle-core / candle-nn /
candle-kernels 0.9.1, gemm 0.17/0.18 -- ``server/rust/moshi/Cargo.lock:539-620,1853-1875``); no
cargo/rustc is available here, so it cannot be compiled into ``oracle/_ref``.

Parity pinning status (see DESIGN.md "Oracle"):
  * PINNED by the reference's own tests: the ring index/mask table
    (``kv_cache.rs:338-405`` -> ``tests/test_oracle_kv_cache.py``) and streaming == one-shot for
    conv1d / conv-transpose1d over the reference's grid (``conv.rs:698-723`` ->
    ``tests/test_oracle_conv.py``).
  * Cross-checked against the independent HF ``transformers`` Mimi implementation with shared random
    weights (``tests/test_oracle_hf_mimi.py``) where that package is importable.
  * PARITY UNPINNED for everything the reference has no tests or golden vectors for: Mimi code
    indices, RVQ, LM logits, ASR transcripts.  For those the oracle is a line-by-line restatement,
    each function citing the file:line it follows; un-vendored candle op semantics it assumes are
    listed in SURVEY.md Appendix C.
"""
