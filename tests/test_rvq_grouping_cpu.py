"""The grouped RVQ encode of csrc/rvq_gram.cu restated in numpy f32 -- scores of a whole group from the group's entry
residual, corrections read from Gram tables E_p E_q^T (double accumulation, rounded to f32), first-index argmin, the
reference's one-subtraction-per-codebook residual update -- must choose the same codes as the oracle's sequential
quantiser (quantization.rs:219-229) at Mimi's sizes.  This pins the algebra and the rounding argument of DESIGN.md
section 5 without a GPU; the GPU parity tests (tests/test_gpu_mimi.py) pin the kernel."""
import numpy as np
import pytest
import torch

from oracle.mimi import ResidualVectorQuantizer


def _weights(n_q, bins, qd, d, seed):
    g = torch.Generator().manual_seed(seed)
    w = {"q.input_proj.weight": torch.randn(qd, d, 1, generator=g) / d ** 0.5,
         "q.output_proj.weight": torch.randn(d, qd, 1, generator=g) / qd ** 0.5}
    for i in range(n_q):
        # codebooks whose scale shrinks level by level, like a trained residual quantiser's
        w[f"q.vq.layers.{i}._codebook.embedding_sum"] = torch.randn(bins, qd, generator=g) * (0.9 ** i)
        w[f"q.vq.layers.{i}._codebook.cluster_usage"] = torch.ones(bins)
    return w


def grouped_encode(rvq: ResidualVectorQuantizer, xs: torch.Tensor, group: int) -> np.ndarray:
    """xs [B, C, 1] -> codes [B, n_q]; f32 throughout, in the order the kernels use."""
    emb = [l.embedding.numpy().astype(np.float32) for l in rvq.layers]
    c2 = [l.c2.numpy().astype(np.float32) for l in rvq.layers]
    r = torch.nn.functional.conv1d(xs, rvq.input_proj)[:, :, 0].numpy().astype(np.float32)      # [B, qd]
    n_q, B = len(emb), r.shape[0]
    codes = np.zeros((B, n_q), np.int64)
    for q0 in range(0, n_q, group):
        g = min(group, n_q - q0)
        # one GEMM for the group: |e|^2/2 - r_q0 . e for every codebook of the group
        base = [c2[q0 + i][None, :] - r @ emb[q0 + i].T for i in range(g)]
        gram = {(p, q): (emb[q0 + p].astype(np.float64) @ emb[q0 + q].astype(np.float64).T).astype(np.float32)
                for p in range(g) for q in range(p + 1, g)}
        for q in range(g):
            v = base[q].copy()
            for p in range(q):                                   # ascending p, like rvq_chain_kernel
                v = v + gram[(p, q)][codes[:, q0 + p]]
            codes[:, q0 + q] = v.argmin(axis=1)                  # numpy's argmin returns the first minimum
        for q in range(g):                                       # residual: one f32 subtraction per codebook, in order
            r = r - emb[q0 + q][codes[:, q0 + q]]
    return codes


@pytest.mark.parametrize("group", [1, 3, 8])
def test_grouped_rvq_chooses_the_oracles_codes(group):
    n_q, bins, qd, d, B = 31, 2048, 256, 512, 96
    rvq = ResidualVectorQuantizer(_weights(n_q, bins, qd, d, seed=3), "q", n_q)
    xs = torch.randn(B, d, 1, generator=torch.Generator().manual_seed(group))
    want = rvq.encode(xs)[:, :, 0].numpy()
    got = grouped_encode(rvq, xs, group)
    same_rows = int((got == want).all(axis=1).sum())
    # group 1 is the reference's own association; larger groups only change how r_q . e_n is rounded
    assert same_rows >= (B if group == 1 else int(0.995 * B)), f"{same_rows}/{B} rows identical"
    assert (got[:, 0] == want[:, 0]).all()


def test_gram_identity_in_double():
    """|e_n|^2/2 - r_q . e_n == (|e_n|^2/2 - r_q0 . e_n) + sum_p e_p[code_p] . e_n, to double rounding."""
    rng = np.random.default_rng(0)
    E = rng.standard_normal((4, 64, 16))
    r0 = rng.standard_normal(16)
    codes = [5, 17, 3]
    r = r0 - sum(E[p][codes[p]] for p in range(3))
    c2 = (E[3] * E[3]).sum(-1) / 2
    lhs = c2 - E[3] @ r
    rhs = (c2 - E[3] @ r0) + sum(E[3] @ E[p][codes[p]] for p in range(3))
    assert np.allclose(lhs, rhs, rtol=0, atol=1e-12)
