"""Shared helpers for the test-suite (fixture loading, tolerant comparisons)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL_TOL = 1e-5   # BASELINE.json: quantize / diff / EMA buffers within 1e-5 relative in fp32


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def rel_err(a, b):
    """max |a-b| relative to the scale of b (max-norm): robust for buffers that mix 1e5 and 1e-2."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-30)
    return float(np.max(np.abs(a - b)) / denom) if b.size else 0.0


def col_rel_err(a, b):
    """per-codebook-column relative error for [D,K] buffers (each code judged on its own scale)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    num = np.max(np.abs(a - b), axis=0)
    den = np.maximum(np.max(np.abs(b), axis=0), 1e-30)
    return float(np.max(num / den))


# ---------------------------------------------------------------------------------------------
# element-wise comparisons (round 2).  BASELINE.json: "quantize, diff and the EMA buffers must match within
# 1e-5 relative in fp32".  An element of an fp32 SUM can cancel to ~0, so "relative" is taken element by element
# against the magnitude of the terms that element sums (the standard forward-error scale of a summation) -- never
# against the largest element of the whole buffer.
# ---------------------------------------------------------------------------------------------
def scaled_err(a, b, scale):
    """max_ij |a_ij - b_ij| / scale_ij (float64; `scale` broadcastable, strictly positive where a != b)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    s = np.broadcast_to(np.asarray(scale, dtype=np.float64), a.shape)
    d = np.abs(a - b)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(d == 0.0, 0.0, d / s)
    return float(np.nanmax(r)) if not np.isnan(r).all() else float("nan")


def elem_rel_err(a, b, floor=1e-30):
    """max_ij |a_ij - b_ij| / max(|b_ij|, floor_ij): per-element relative error with an absolute floor."""
    b64 = np.asarray(b, dtype=np.float64)
    return scaled_err(a, b, np.maximum(np.abs(b64), floor))


def ema_scales(x_flat, ind, embed_avg_before, cluster_size_after, decay, eps):
    """Element-wise magnitude of the terms every EMA-buffer element sums (vqvae.py:61-70), float64:
         embed_avg[d,k] = decay*embed_avg_old[d,k] + (1-decay) * sum_{i: z_i=k} x[i,d]  -> decay*|old| + (1-decay)*sum|x|
         embed[d,k]     = embed_avg[d,k] / cs_hat[k]                                     -> the same / cs_hat[k]
    Returns (scale_embed_avg [D,K], scale_embed [D,K])."""
    x = np.asarray(x_flat, dtype=np.float64)
    ind = np.asarray(ind).reshape(-1)
    K = embed_avg_before.shape[1]
    sabs = np.zeros((K, x.shape[1]))
    np.add.at(sabs, ind, np.abs(x))
    scale_avg = decay * np.abs(np.asarray(embed_avg_before, dtype=np.float64)) + (1.0 - decay) * sabs.T
    cs = np.asarray(cluster_size_after, dtype=np.float64)
    n = cs.sum()
    cs_hat = (cs + eps) / (n + K * eps) * n
    scale_avg = np.maximum(scale_avg, 1e-30)
    return scale_avg, scale_avg / np.maximum(cs_hat[None, :], 1e-30)


def near_tie_columns(ind_a, ind_b, n_embed):
    """Boolean keep-mask over the codes: False for every code a tolerated near-tie row touches in either index vector (one
    moved vector changes those EMA columns by ~1e-3); plus the differing rows."""
    a, b = np.asarray(ind_a).reshape(-1), np.asarray(ind_b).reshape(-1)
    keep = np.ones(n_embed, dtype=bool)
    rows = np.nonzero(a != b)[0]
    keep[a[rows]] = False
    keep[b[rows]] = False
    return keep, rows


def check_outputs_np(tag, x_np, state_before, outs, want, got_state, want_state, train, decay=0.99, eps=1e-5):
    """Element-wise 1e-5 check of one forward.  outs / want = (quantize, diff, ind) as numpy; got_state / want_state =
    (cluster_size, embed_avg, embed) AFTER the call (training only); state_before = dict of the buffers before it.
    Index rule: exact except float64 near-ties (< 1e-6 relative); the EMA columns of codes touched by a tolerated
    near-tie row are excluded.  Returns the number of tolerated near-tie rows."""
    from oracle.quantize_oracle import tie_tolerant_index_mismatches
    quant, diff, ind_np = outs
    wq, wd, wi = want
    embed_before = state_before["embed"]
    D, K = embed_before.shape
    assert ind_np.dtype == np.int64 and tuple(ind_np.shape) == tuple(x_np.shape[:-1])
    ndiff, nbad, bad = tie_tolerant_index_mismatches(x_np, embed_before, ind_np, wi)
    assert nbad == 0, f"{tag}: {nbad} index mismatches beyond fp32 near-ties (rows {bad[:8]})"
    keep, rows = near_tie_columns(ind_np, wi, K)
    flat = x_np.reshape(-1, D)
    same = np.ones(flat.shape[0], dtype=bool)
    same[rows] = False
    # quantize must be the gather of the chosen code from the PRE-update codebook (vqvae.py:52,73), element by element
    codes = embed_before.T[ind_np.reshape(-1)]
    qn = np.asarray(quant).reshape(-1, D)
    qscale = np.maximum(np.abs(codes), np.abs(flat)).astype(np.float64) + 1e-30
    assert scaled_err(qn, flat + (codes - flat), qscale) <= REL_TOL, f"{tag}: quantize is not the gather of the chosen codes"
    assert scaled_err(qn[same], np.asarray(wq).reshape(-1, D)[same], qscale[same]) <= REL_TOL, f"{tag}: quantize"
    slack = 0.0
    if ndiff:
        e64 = embed_before.astype(np.float64)
        xr = flat[rows].astype(np.float64)
        slack = float(np.abs(((xr - e64[:, ind_np.reshape(-1)[rows]].T) ** 2).sum(1)
                             - ((xr - e64[:, np.asarray(wi).reshape(-1)[rows]].T) ** 2).sum(1)).sum()) / flat.size
    assert abs(float(diff) - float(wd)) <= REL_TOL * abs(float(wd)) + slack + 1e-30, f"{tag}: diff {float(diff)!r} vs {float(wd)!r}"
    if train:
        g_cs, g_avg, g_emb = (np.asarray(v) for v in got_state)
        w_cs, w_avg, w_emb = (np.asarray(v) for v in want_state)
        s_avg, s_emb = ema_scales(flat, np.asarray(wi).reshape(-1), state_before["embed_avg"], w_cs, decay, eps)
        errs = {"cluster_size": elem_rel_err(g_cs[keep], w_cs[keep]),
                "embed_avg": scaled_err(g_avg[:, keep], w_avg[:, keep], s_avg[:, keep]),
                "embed": scaled_err(g_emb[:, keep], w_emb[:, keep], s_emb[:, keep])}
        for name, err in errs.items():
            assert err <= REL_TOL, f"{tag}: {name} off by {err:.3e} element-wise (tolerance {REL_TOL:g})"
    return ndiff
