"""Comparison harness: a drop-in `Quantize` against the reference's own `Quantize` (oracle/_ref, unmodified) on the same
device, multi-step, at any size.  Device-agnostic torch code: the `-m gpu` tests run both modules on cuda:0 (the
reference with TF32 off, so vqvae.py:46 is a true fp32 product); a CPU test runs the harness reference-vs-oracle to keep
the harness itself honest.

Tolerances (BASELINE.json north_star):
  * embed_ind: exact, except rows where the float64 distances of the two chosen codes differ by < 1e-6 relative to the
    magnitude of the terms the fp32 expression sums (||x||^2 + ||e||^2);
  * quantize, diff, cluster_size, embed_avg, embed: 1e-5 relative, ELEMENT BY ELEMENT, each element judged against the
    magnitude of the terms it sums (tests/helpers.py `ema_scales`); the EMA-buffer columns of the (at most a handful of)
    codes touched by a tolerated near-tie row are excluded, because one moved vector changes those columns by ~1e-3
    (SURVEY section 7), and are reported.
After every step the candidate's buffers are overwritten with the reference's, so a tolerated near-tie never forks the
trajectory and no test needs to skip.
"""
from __future__ import annotations

import json
import os

import torch
import torch.nn.functional as F

TOL = 1e-5
TIE_REL = 1e-6
REPORT = []          # one dict per compared step; dumped by the tests into gpurun_out/parity_report.json


def fp32_reference_backends():
    """The reference must compute vqvae.py:46 in true fp32 (torch's default; set explicitly, other tests may differ)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")


def sync_state(dst, src):
    """dst's three buffers <- src's (follow the reference trajectory)."""
    with torch.no_grad():
        dst.embed.data.copy_(src.embed.data)
        dst.cluster_size.data.copy_(src.cluster_size.data)
        dst.embed_avg.data.copy_(src.embed_avg.data)


def snapshot(mod):
    return {"embed": mod.embed.detach().clone(), "cluster_size": mod.cluster_size.detach().clone(),
            "embed_avg": mod.embed_avg.detach().clone()}


def index_mismatches(x_flat, embed_before, ia, ib, rel=TIE_REL):
    """(n_differ, n_bad, rows, codes): rows where the index vectors differ; bad = float64 distance gap of the two chosen
    codes >= rel * (||x||^2 + max ||e||^2); codes = every code such a row touches in either vector."""
    ia, ib = ia.reshape(-1), ib.reshape(-1)
    rows = torch.nonzero(ia != ib).reshape(-1)
    if rows.numel() == 0:
        return 0, 0, rows, rows
    xr = x_flat[rows].double()
    ea = embed_before[:, ia[rows]].t().double()
    eb = embed_before[:, ib[rows]].t().double()
    da = ((xr - ea) ** 2).sum(1)
    db = ((xr - eb) ** 2).sum(1)
    scale = (xr * xr).sum(1) + torch.maximum((ea * ea).sum(1), (eb * eb).sum(1))
    bad = (da - db).abs() > rel * scale
    codes = torch.unique(torch.cat([ia[rows], ib[rows]]))
    return int(rows.numel()), int(bad.sum()), rows, codes


def ema_scales(x_flat, ind, embed_avg_before, cluster_size_after, decay, eps):
    """torch/float64 twin of tests/helpers.py `ema_scales` (works at N = 4e6 on the GPU)."""
    K = embed_avg_before.shape[1]
    sabs = torch.zeros(K, x_flat.shape[1], dtype=torch.float64, device=x_flat.device)
    sabs.index_add_(0, ind.reshape(-1), x_flat.abs().double())
    scale_avg = decay * embed_avg_before.double().abs() + (1.0 - decay) * sabs.t()
    cs = cluster_size_after.double()
    n = cs.sum()
    cs_hat = (cs + eps) / (n + K * eps) * n
    scale_avg = scale_avg.clamp_min(1e-30)
    return scale_avg, scale_avg / cs_hat.clamp_min(1e-30)[None, :]


def scaled_err(a, b, scale):
    d = (a.double() - b.double()).abs()
    r = torch.where(d == 0, torch.zeros_like(d), d / scale)
    return float(r.max()) if r.numel() else 0.0


@torch.no_grad()
def chunked_reference_forward(ref_q, x, chunk):
    """The reference forward (vqvae.py:42-75) evaluated in row chunks for sweep points whose six [N,K] temporaries do not
    fit: distance / arg-min / gather run through the reference module itself in eval mode per chunk (rows are
    independent); the one-hot statistics use the reference's expressions (:50,55-56) per chunk and add across chunks;
    the EMA update restates :61-70 with the same torch calls on the module's own buffers."""
    was_training = ref_q.training
    flat = x.reshape(-1, ref_q.dim)
    n = flat.shape[0]
    quant = torch.empty_like(flat)
    ind = torch.empty(n, dtype=torch.int64, device=x.device)
    sq = torch.zeros((), dtype=torch.float64, device=x.device)
    onehot_sum = torch.zeros(ref_q.n_embed, dtype=torch.float32, device=x.device)
    embed_sum = torch.zeros(ref_q.dim, ref_q.n_embed, dtype=torch.float32, device=x.device)
    ref_q.eval()
    for s in range(0, n, chunk):
        blk = flat[s:s + chunk]
        q, d, i = ref_q(blk)                                               # vqvae.py:43-52, 72-73 on this chunk
        quant[s:s + chunk] = q
        ind[s:s + chunk] = i
        sq += d.double() * blk.numel()
        if was_training:
            onehot = F.one_hot(i, ref_q.n_embed).type(blk.dtype)           # :50
            onehot_sum += onehot.sum(0)                                    # :55
            embed_sum += blk.transpose(0, 1) @ onehot                      # :56
            del onehot
    ref_q.train(was_training)
    if was_training:
        ref_q.cluster_size.data.mul_(ref_q.decay).add_(onehot_sum, alpha=1 - ref_q.decay)      # :61-63
        ref_q.embed_avg.data.mul_(ref_q.decay).add_(embed_sum, alpha=1 - ref_q.decay)          # :64
        nn_ = ref_q.cluster_size.sum()                                                         # :65
        cluster_size = (ref_q.cluster_size + ref_q.eps) / (nn_ + ref_q.n_embed * ref_q.eps) * nn_   # :66-68
        ref_q.embed.data.copy_(ref_q.embed_avg / cluster_size.unsqueeze(0))                    # :69-70
    diff = (sq / flat.numel()).float()
    return quant.reshape(x.shape), diff, ind.reshape(x.shape[:-1])


def compare_step(tag, ref_q, our_q, x, ref_chunk=None, check_strides=True):
    """One forward of both modules on the same input and the same state; asserts the tolerances above; returns the
    report entry.  Leaves `our_q`'s buffers equal to the reference's."""
    D, K = ref_q.dim, ref_q.n_embed
    train = ref_q.training
    assert our_q.training == train
    before = snapshot(ref_q)
    with torch.no_grad():
        for k, v in snapshot(our_q).items():
            assert torch.equal(v, before[k]), f"{tag}: candidate does not start from the reference state ({k})"
        if ref_chunk:
            rq, rd, ri = chunked_reference_forward(ref_q, x, ref_chunk)
        else:
            rq, rd, ri = ref_q(x)
        oq, od, oi = our_q(x)
    assert oi.dtype == torch.int64 and oi.shape == x.shape[:-1] and oi.is_contiguous()
    assert oq.shape == x.shape and oq.dtype == torch.float32 and od.dim() == 0 and od.dtype == torch.float32
    if check_strides and not ref_chunk:
        # (strides of size-1 dimensions are arbitrary in torch: compared where they address anything)
        assert all(a == b for a, b, n in zip(oq.stride(), rq.stride(), x.shape) if n > 1), \
            f"{tag}: quantize strides {oq.stride()} != reference {rq.stride()}"
    flat = x.reshape(-1, D)
    n = flat.shape[0]
    n_differ, n_bad, rows, codes = index_mismatches(flat, before["embed"], oi, ri)
    assert n_bad == 0, f"{tag}: {n_bad} of {n} rows disagree with the reference beyond an fp32 near-tie (rows {rows[:8].tolist()})"
    entry = {"case": tag, "rows": n, "dim": D, "n_embed": K, "train": bool(train), "index_differ_near_tie": n_differ}
    same = torch.ones(n, dtype=torch.bool, device=x.device)
    same[rows] = False
    # quantize (vqvae.py:73): element-wise, against max(|value|, |x|) -- x + (q - x) rounds at the scale of x
    oqf, rqf = oq.reshape(-1, D)[same], rq.reshape(-1, D)[same]
    entry["quantize"] = scaled_err(oqf, rqf, torch.maximum(rqf.abs(), flat[same].abs()).double().clamp_min(1e-30))
    assert entry["quantize"] <= TOL, f"{tag}: quantize off by {entry['quantize']:.3e} relative"
    # diff (vqvae.py:72): the near-tie rows contribute their own (equal to < 1e-6) distances
    slack = 0.0
    if n_differ:
        xr = flat[rows].double()
        slack = float((((xr - before["embed"][:, oi.reshape(-1)[rows]].t().double()) ** 2).sum(1)
                       - ((xr - before["embed"][:, ri.reshape(-1)[rows]].t().double()) ** 2).sum(1)).abs().sum()) / flat.numel()
    entry["diff"] = abs(float(od) - float(rd)) / max(abs(float(rd)), 1e-30)
    assert abs(float(od) - float(rd)) <= TOL * abs(float(rd)) + slack, f"{tag}: diff {float(od)!r} vs reference {float(rd)!r}"
    if train:
        keep = torch.ones(K, dtype=torch.bool, device=x.device)
        keep[codes] = False
        entry["ema_columns_excluded"] = int(codes.numel())
        s_avg, s_emb = ema_scales(flat, ri, before["embed_avg"], ref_q.cluster_size, ref_q.decay, ref_q.eps)
        entry["cluster_size"] = scaled_err(our_q.cluster_size[keep], ref_q.cluster_size[keep],
                                           ref_q.cluster_size[keep].double().abs().clamp_min(1e-30))
        entry["embed_avg"] = scaled_err(our_q.embed_avg[:, keep], ref_q.embed_avg[:, keep], s_avg[:, keep])
        entry["embed"] = scaled_err(our_q.embed[:, keep], ref_q.embed[:, keep], s_emb[:, keep])
        for name in ("cluster_size", "embed_avg", "embed"):
            assert entry[name] <= TOL, f"{tag}: {name} off by {entry[name]:.3e} (element-wise, tolerance {TOL:g})"
        sync_state(our_q, ref_q)
    else:
        for k, v in snapshot(our_q).items():
            assert torch.equal(v, before[k]), f"{tag}: eval mode touched {k} (vqvae.py:54)"
    REPORT.append(entry)
    return entry


def dump_report(name="parity_report.json"):
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, name), "w") as f:
            json.dump(REPORT, f, indent=1)
    except OSError:
        pass


def make_inputs(kind, shape, embed, seed, device, permuted=False):
    """SURVEY 8d value distributions: 'randn' (reference-init regime), 'clustered' (embed[:, randint(K)] + 0.1 N(0,1) against
    the CURRENT codebook; dead ~1e5 codes are skipped so the rows stay finite-scale).  `permuted`: the NCHW-physical
    permute(0,2,3,1) view VQVAE.encode passes (vqvae.py:227,235)."""
    g = torch.Generator(device=device).manual_seed(seed)
    D = shape[-1]
    n = 1
    for s in shape[:-1]:
        n *= s
    if kind == "randn":
        x = torch.randn(n, D, device=device, generator=g)
    elif kind == "clustered":
        norms = embed.pow(2).sum(0)
        live = torch.nonzero(norms < 100.0 * norms.min().clamp_min(1e-12)).reshape(-1)
        pick = live[torch.randint(0, live.numel(), (n,), device=device, generator=g)]
        x = embed.t()[pick] + 0.1 * torch.randn(n, D, device=device, generator=g)
    else:
        raise ValueError(kind)
    x = x.reshape(shape).contiguous()
    if permuted:
        assert len(shape) == 4
        x = x.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    return x
