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
  * CROSS-CHECKED against independent implementations of the same models, the HF ``transformers``
    ``MimiModel`` and ``KyutaiSpeechToTextModel``, loaded with the SAME random weights:
    encoder latents + RVQ codes and decoded audio (``tests/test_oracle_hf_mimi.py``), LM hidden
    states + greedy tokens (``tests/test_oracle_hf_lm.py``).  This pins the architecture (layer
    order, strides, paddings, ELU / LayerScale / gating placement, RoPE pairing, RVQ residual
    updates); it does not pin candle's rounding.
  * PARITY UNPINNED against the reference's own outputs: the reference has no tests, golden
    vectors or fixtures for Mimi codes, RVQ indices, LM logits or transcripts, and it cannot be
    built here to generate any.  Un-vendored candle op semantics the oracle assumes are listed in
    SURVEY.md Appendix C.
"""
