"""GPU tests of the WIDE tcgen05 engine (csrc/tc_wide_kernel.cuh): dim = 128 / 256 -- the D axis of BASELINE.json's cfg-5
sweep and the D = 256 quantizers of vqvae_deep.py:252,257.  Same bar as the D = 64 engine: the tensor-core scores are
certified lower bounds of the float64 distances, the indices after the exact fix-up are the exact-arithmetic arg-min, and
outputs / EMA buffers match the exact fp32 SIMT engine and the oracle."""
import numpy as np
import pytest
import torch

import vq_vae_2_pytorch_b200 as vq
from vq_vae_2_pytorch_b200 import _native
from helpers import REL_TOL, col_rel_err, rel_err
from oracle.quantize_oracle import QuantizeOracle, distances_f64, tie_tolerant_index_mismatches
from test_gpu_tc import tc_scores

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# (D, K): one launch (resident image) / two slices (x converted per slice) / four slices (x converted once, operand stages streamed)
SHAPES = [(128, 512), (128, 256), (256, 256), (256, 512), (128, 1024), (128, 2048), (256, 1024)]


def test_wide_shapes_are_reported_as_tensor_core_shapes():
    lib = _native.load()
    x = torch.zeros(256, 256, device=DEV)
    for d, k in SHAPES:
        assert lib.vqb200_tc_supported(_native.ptr(x), 256, d, k, 256, 0, d, 1) == 1
    assert lib.vqb200_tc_supported(_native.ptr(x), 256, 192, 512, 256, 0, 192, 1) == 0
    assert lib.vqb200_tc_supported(_native.ptr(x), 256, 128, 384, 256, 0, 128, 1) == 0


@pytest.mark.parametrize("D,K", SHAPES)
@pytest.mark.parametrize("cb", ["randn", "dead", "small"])
def test_wide_scores_are_certified_lower_bounds(D, K, cb):
    rng = np.random.default_rng(3)
    embed = rng.standard_normal((D, K)).astype(np.float32)
    if cb == "dead":
        embed[:, 60:] *= 1.0e5                 # collapsed regime
    if cb == "small":
        embed *= 1e-2
    n = 1000                                   # ragged: 7 full tiles + 104 rows
    x = ((1e-2 if cb == "small" else 1.0) * rng.standard_normal((n, D))).astype(np.float32)
    x[:64] = embed[:, rng.integers(0, 60, 64)].T
    scores, ind, flagged = tc_scores(x, embed)
    assert not np.isnan(scores).any(), "tensor-core scores were not written for every (row, code)"
    cA, cB = 7.9e-3, 4.0e-6 * (D // 64)        # plain-bf16 filter; accumulation term scales with the MMAs accumulated
    d64 = distances_f64(x, embed)
    xx = (x.astype(np.float64) ** 2).sum(1, keepdims=True)
    ee = (embed.astype(np.float64) ** 2).sum(0, keepdims=True)
    target = d64 - xx + xx * (1.0 + 2.0 ** -9)
    ebound = cA * np.sqrt(xx) * np.sqrt(ee) + cB * (ee + xx)
    tiny = 1e-6 * (xx + ee)
    err = scores.astype(np.float64) - target
    assert (err <= tiny).all(), f"score above the true distance by {err.max():.3e} (bound violated)"
    assert (err >= -(2.2 * ebound + tiny)).all(), "score is looser than the documented bound"
    raw = scores.astype(np.float64) + (cA * np.sqrt(xx) * np.sqrt(ee) + cB * ee) - target
    print(f"[tcw D={D} K={K} {cb}] flagged={flagged}/{n} max|filter err|/(|x||e|)={(np.abs(raw) / (np.sqrt(xx) * np.sqrt(ee) + 1e-30)).max():.3e}"
          f" max|err|/(xx+ee)={(np.abs(raw) / (xx + ee)).max():.3e}")
    o = QuantizeOracle(D, K, embed=embed)
    o.training = False
    _, _, io = o.forward(x)
    _, nbad, _ = tie_tolerant_index_mismatches(x, embed, ind, io)
    assert nbad == 0
    assert (ind >= 0).all() and (ind < K).all()


@pytest.mark.parametrize("D,K", SHAPES)
def test_wide_engine_matches_simt_engine_and_oracle(D, K):
    torch.manual_seed(7)
    N = 128 * 41 + 77
    a = vq.Quantize(D, K, engine="tcgen05").to(DEV).train()
    b = vq.Quantize(D, K, engine="simt").to(DEV).train()
    b.load_state_dict(a.state_dict())
    embed0 = a.embed.clone()
    pick = torch.randint(0, K, (N,), device=DEV)
    x = torch.cat([embed0.t()[pick[: N // 2]] + 0.2 * torch.randn(N // 2, D, device=DEV),
                   torch.randn(N - N // 2, D, device=DEV)]).contiguous()       # half clustered, half N(0,1) (many near-ties)
    o = QuantizeOracle(D, K, embed=embed0.cpu().numpy())
    for step in range(2):                      # second step: codebook with dead ~1e5-magnitude codes
        embed_before = b.embed.cpu().numpy().copy()
        qa, da, ia = a(x)
        qb, db, ib = b(x)
        qo, do, io = o.forward(x.cpu().numpy())
        _, nbad, _ = tie_tolerant_index_mismatches(x.cpu().numpy(), embed_before, ia.cpu().numpy(), ib.cpu().numpy())
        assert nbad == 0
        _, nbad, _ = tie_tolerant_index_mismatches(x.cpu().numpy(), embed_before, ia.cpu().numpy(), io)
        assert nbad == 0
        if int((ia != ib).sum()) == 0:
            assert torch.equal(qa, qb)
            assert abs(float(da) - float(db)) <= 1e-5 * abs(float(db))
            assert col_rel_err(a.embed_avg.cpu().numpy(), b.embed_avg.cpu().numpy()) <= REL_TOL
            assert torch.allclose(a.cluster_size, b.cluster_size, rtol=1e-5, atol=1e-7)
        if int((ia.cpu().numpy() != io).sum()) == 0:
            assert rel_err(qa.cpu().numpy(), qo) <= REL_TOL
            assert abs(float(da) - float(do)) <= REL_TOL * abs(float(do))
        b.load_state_dict(a.state_dict())
        o.embed, o.cluster_size, o.embed_avg = (t.cpu().numpy().copy() for t in (a.embed, a.cluster_size, a.embed_avg))


@pytest.mark.parametrize("D,K", [(128, 512), (256, 512)])
def test_wide_engine_eval_many_trips_and_clustered_rows(D, K):
    """More tiles than CTAs (several trips per CTA, ragged tail); clustered rows must return to their generating code."""
    torch.manual_seed(11)
    N = 128 * 148 * 2 + 128 * 9 + 5
    q = vq.Quantize(D, K, engine="tcgen05").to(DEV).eval()
    pick = torch.randint(0, K, (N,), device=DEV)
    x = (q.embed.t()[pick] + 0.05 * torch.randn(N, D, device=DEV)).contiguous()
    quant, diff, ind = q(x)
    assert torch.equal(ind, pick)
    assert torch.allclose(quant, q.embed.t()[pick], rtol=1e-5, atol=1e-6)
    ref = ((q.embed.t()[pick] - x) ** 2).mean()
    assert abs(float(diff) - float(ref)) <= 1e-5 * float(ref)


def test_wide_engine_on_deep_fork_shapes_with_permuted_input():
    """vqvae_deep.py:288-299 passes permute(0,2,3,1) views of [B,256,36,18] / [B,256,18,9]: not a tensor-core layout, the
    module re-packs to dense rows around the wide engine; `quantize` keeps the input's strides (vqvae.py:73)."""
    torch.manual_seed(13)
    for (B, H, W) in [(8, 36, 18), (8, 18, 9)]:
        a = vq.Quantize(256, 512).to(DEV).train()
        b = vq.Quantize(256, 512, engine="simt").to(DEV).train()
        b.load_state_dict(a.state_dict())
        x = torch.randn(B, 256, H, W, device=DEV).permute(0, 2, 3, 1)
        embed_before = a.embed.cpu().numpy().copy()
        qa, da, ia = a(x)
        qb, db, ib = b(x)
        assert qa.stride() == x.stride() and ia.shape == (B, H, W)
        flat = x.reshape(-1, 256).cpu().numpy()
        _, nbad, _ = tie_tolerant_index_mismatches(flat, embed_before, ia.reshape(-1).cpu().numpy(), ib.reshape(-1).cpu().numpy())
        assert nbad == 0
        if int((ia != ib).sum()) == 0:
            assert torch.equal(qa, qb)
            assert col_rel_err(a.embed_avg.cpu().numpy(), b.embed_avg.cpu().numpy()) <= REL_TOL


@pytest.mark.parametrize("D,K", [(128, 512), (256, 512), (256, 1024)])
@pytest.mark.parametrize("N", [1, 63, 65, 129])
def test_wide_engine_tiny_inputs(D, K, N):
    """Fewer rows than one TMA box / one tile (the tensor map's row extent is smaller than its box)."""
    torch.manual_seed(17)
    a = vq.Quantize(D, K, engine="tcgen05").to(DEV).eval()
    b = vq.Quantize(D, K, engine="simt").to(DEV).eval()
    b.load_state_dict(a.state_dict())
    pick = torch.randint(0, K, (N,), device=DEV)
    x = (a.embed.t()[pick] + 0.1 * torch.randn(N, D, device=DEV)).contiguous()
    qa, da, ia = a(x)
    qb, db, ib = b(x)
    assert torch.equal(ia, pick) and torch.equal(ib, pick)
    assert torch.equal(qa, qb)
    assert abs(float(da) - float(db)) <= 1e-6 * abs(float(db))
