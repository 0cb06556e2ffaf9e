"""CPU oracle for the VQ-VAE-2 `Quantize` hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the algorithm of the reference module
`Quantize` (/root/reference/vqvae.py:28-78).  It is the *checker* for the CUDA
path: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it.  The product package
(`vq_vae_2_pytorch_b200`) never imports anything from `oracle/`.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4 / 8c), so the oracle is pinned against outputs of the
reference module itself, executed in the build container on seeded inputs by
`tests/golden/make_golden.py`; the resulting fixtures live in `tests/golden/`
and `tests/test_oracle_golden.py` checks this file against every one of them.

Arithmetic: every step is evaluated in fp32, in the order the reference
evaluates it (each function cites the reference lines it follows).  The only
third-party arithmetic on the reference's path is the BLAS sgemm behind the two
`@` products (MKL on CPU, cuBLAS on GPU, torch 2.11); here it is numpy's
OpenBLAS sgemm, i.e. the summation order inside the contraction is, as in the
reference, unspecified.  `distances_f64` gives the float64 ground truth that the
tests use to classify fp32 near-ties.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def distances_f32(flatten: np.ndarray, embed: np.ndarray) -> np.ndarray:
    """vqvae.py:44-48.  ||x||^2 - (2x)@E + ||e||^2, evaluated left to right in fp32.

    `2 * flatten @ self.embed` parses as `(2 * flatten) @ embed` (vqvae.py:46).
    """
    xx = np.sum(flatten * flatten, axis=1, keepdims=True, dtype=F32)       # :45
    xe = (F32(2.0) * flatten) @ embed                                      # :46
    ee = np.sum(embed * embed, axis=0, keepdims=True, dtype=F32)           # :47
    return (xx - xe) + ee


def distances_f64(flatten: np.ndarray, embed: np.ndarray) -> np.ndarray:
    """Float64 evaluation of the same expression; near-tie classifier for the tests."""
    x = flatten.astype(np.float64)
    e = embed.astype(np.float64)
    return (x * x).sum(1, keepdims=True) - 2.0 * (x @ e) + (e * e).sum(0, keepdims=True)


def nearest_code(dist: np.ndarray) -> np.ndarray:
    """vqvae.py:49.  `(-dist).max(1)` index; first (lowest) index on exact ties
    (numpy argmax and torch CPU max agree on that)."""
    return np.argmax(-dist, axis=1).astype(np.int64)


def embed_code(embed: np.ndarray, embed_id: np.ndarray) -> np.ndarray:
    """vqvae.py:77-78.  F.embedding(embed_id, embed.T) -> [..., dim]."""
    return np.ascontiguousarray(embed.T)[embed_id]


def code_statistics(flatten: np.ndarray, embed_ind: np.ndarray, n_embed: int):
    """vqvae.py:50,55-56.  one-hot column sums and flatten^T @ onehot (fp32)."""
    onehot = np.zeros((flatten.shape[0], n_embed), dtype=F32)              # :50
    onehot[np.arange(flatten.shape[0]), embed_ind] = F32(1.0)
    onehot_sum = onehot.sum(0, dtype=F32)                                  # :55
    embed_sum = flatten.T @ onehot                                         # :56
    return onehot_sum, embed_sum


def ema_update(cluster_size, embed_avg, onehot_sum, embed_sum, decay, eps, n_embed):
    """vqvae.py:61-70.  Returns the three new buffers (cluster_size, embed_avg, embed).

    `1 - decay` is evaluated in Python double and then applied as an fp32 scalar
    (torch `add_(t, alpha=1 - decay)`), vqvae.py:61-64.
    """
    d = F32(decay)
    a = F32(1 - decay)
    cluster_size = cluster_size * d + onehot_sum * a                       # :61-63
    embed_avg = embed_avg * d + embed_sum * a                              # :64
    n = cluster_size.sum(dtype=F32)                                        # :65
    cs = (cluster_size + F32(eps)) / (n + F32(n_embed * eps)) * n          # :66-68
    embed = embed_avg / cs[None, :]                                        # :69-70
    return cluster_size.astype(F32), embed_avg.astype(F32), embed.astype(F32)


class QuantizeOracle:
    """Stateful restatement of `Quantize` (vqvae.py:28-78) on numpy fp32 arrays.

    `all_reduce` (optional callable on an fp32 array, in place / returning the
    reduced array) stands for `dist_fn.all_reduce` (vqvae.py:58-59 ->
    distributed/distributed.py:64-72).
    """

    def __init__(self, dim, n_embed, decay=0.99, eps=1e-5, embed=None, seed=0):
        self.dim, self.n_embed, self.decay, self.eps = dim, n_embed, decay, eps
        if embed is None:                                                  # :37 (randn)
            embed = np.random.default_rng(seed).standard_normal((dim, n_embed))
        self.embed = np.array(embed, dtype=F32, order="C")                 # :38
        self.cluster_size = np.zeros(n_embed, dtype=F32)                   # :39
        self.embed_avg = self.embed.copy()                                 # :40
        self.training = True

    def state(self):
        return {"embed": self.embed.copy(), "cluster_size": self.cluster_size.copy(),
                "embed_avg": self.embed_avg.copy()}

    def load(self, embed, cluster_size, embed_avg):
        self.embed = np.array(embed, dtype=F32)
        self.cluster_size = np.array(cluster_size, dtype=F32)
        self.embed_avg = np.array(embed_avg, dtype=F32)

    def forward(self, x: np.ndarray, all_reduce=None, row_chunk: int | None = None):
        """vqvae.py:42-75.  Returns (quantize, diff, embed_ind).

        `row_chunk` evaluates distance/argmin/statistics in row blocks (rows are
        independent; counts and sums add across blocks) so sweep points whose
        [N,K] temporaries do not fit can still be checked.
        """
        if x.dtype != F32:
            raise TypeError("Quantize expects float32 input")              # '@' dtype error
        if x.shape[-1] != self.dim:
            raise ValueError("last dimension must equal dim")              # reshape error
        flatten = np.ascontiguousarray(x).reshape(-1, self.dim)            # :43
        n = flatten.shape[0]
        step = n if not row_chunk else row_chunk
        ind = np.empty(n, dtype=np.int64)
        onehot_sum = np.zeros(self.n_embed, dtype=F32)
        embed_sum = np.zeros((self.dim, self.n_embed), dtype=F32)
        for s in range(0, max(n, 1), max(step, 1)):
            blk = flatten[s:s + step]
            ind[s:s + step] = nearest_code(distances_f32(blk, self.embed))  # :44-49
            if self.training:
                c, es = code_statistics(blk, ind[s:s + step], self.n_embed)  # :50,55-56
                onehot_sum += c
                embed_sum += es
        embed_ind = ind.reshape(x.shape[:-1])                              # :51
        quantize = embed_code(self.embed, embed_ind)                       # :52 (pre-update codebook)
        if self.training:                                                  # :54
            if all_reduce is not None:                                     # :58-59
                onehot_sum = all_reduce(onehot_sum)
                embed_sum = all_reduce(embed_sum)
            self.cluster_size, self.embed_avg, self.embed = ema_update(
                self.cluster_size, self.embed_avg, onehot_sum, embed_sum,
                self.decay, self.eps, self.n_embed)                        # :61-70
        d = quantize - x
        diff = np.mean(d * d, dtype=F32)                                   # :72
        quantize = x + (quantize - x)                                      # :73 (value of the STE expr)
        return quantize.astype(F32), F32(diff), embed_ind                  # :75

    def backward(self, x, quantize_codes, grad_quantize, grad_diff):
        """Gradient implied by vqvae.py:72-73: identity through the STE plus
        2(x - q)/(N*D) from `diff`; `quantize_codes` is embed_code(embed_ind)."""
        scale = F32(2.0 / x.size) * F32(grad_diff)
        return (grad_quantize + scale * (x - quantize_codes)).astype(F32)


def tie_tolerant_index_mismatches(x, embed, ind_a, ind_b, rel=1e-6):
    """Rows where two index vectors differ by MORE than an fp32 near-tie.

    A differing row is tolerated when the float64 distances of the two chosen
    codes differ by less than `rel` times the magnitude of the terms the fp32
    expression sums (||x||^2 + ||e||^2) -- BASELINE.json's "distance gap < 1e-6
    relative".  Returns (n_differ, n_bad, bad_rows).
    """
    flat = x.reshape(-1, x.shape[-1]).astype(np.float64)
    a = np.asarray(ind_a).reshape(-1)
    b = np.asarray(ind_b).reshape(-1)
    rows = np.nonzero(a != b)[0]
    if rows.size == 0:
        return 0, 0, rows
    e = embed.astype(np.float64)
    xr = flat[rows]
    ea, eb = e[:, a[rows]].T, e[:, b[rows]].T
    da = ((xr - ea) ** 2).sum(1)
    db = ((xr - eb) ** 2).sum(1)
    scale = (xr * xr).sum(1) + np.maximum((ea * ea).sum(1), (eb * eb).sum(1))
    bad = np.abs(da - db) > rel * scale
    return int(rows.size), int(bad.sum()), rows[bad]
