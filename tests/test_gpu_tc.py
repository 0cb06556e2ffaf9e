"""GPU tests specific to the tcgen05 engine: the tensor-core scores must be certified lower bounds of the
float64 distances (so that the certificate in the epilogue is rigorous), the flagged-row fix-up must give
exact-arithmetic indices, and the fused outputs must match the SIMT engine bit-for-bit where indices agree."""
import ctypes as C

import numpy as np
import pytest
import torch

import vq_vae_2_pytorch_b200 as vq
from vq_vae_2_pytorch_b200 import _native
from helpers import REL_TOL, col_rel_err, rel_err
from oracle.quantize_oracle import QuantizeOracle, distances_f64, tie_tolerant_index_mismatches

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def tc_scores(x, embed):
    """Run vqb200_debug_tc_scores; returns (scores [N,K], embed_ind [N], flagged_count)."""
    lib = _native.load()
    n, d = x.shape
    k = embed.shape[1]
    xd = torch.from_numpy(x).to(DEV)
    ed = torch.from_numpy(embed).to(DEV)
    image = torch.empty(lib.vqb200_codebook_bytes(d, k), dtype=torch.uint8, device=DEV)
    scratch = torch.empty(lib.vqb200_forward_scratch_bytes(n, d, k), dtype=torch.uint8, device=DEV)
    scores = torch.full((n, k), float("nan"), device=DEV)
    ind = torch.full((n,), -1, dtype=torch.int64, device=DEV)
    flagged = torch.zeros(1, dtype=torch.int32, device=DEV)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _native.check(lib.vqb200_codebook_prepare(_native.ptr(ed), d, k, _native.ptr(image), st), "prepare")
    _native.check(lib.vqb200_debug_tc_scores(_native.ptr(xd), n, d, k, _native.ptr(image), _native.ptr(ind),
                                             _native.ptr(scores), _native.ptr(flagged), _native.ptr(scratch), st), "scores")
    torch.cuda.synchronize()
    return scores.cpu().numpy(), ind.cpu().numpy(), int(flagged.item())


def bound_constants():
    split = _native.load().vqb200_tc_split()
    cA = 3.0517578125e-5 if split == 3 else 7.9e-3
    return split, cA, 4.0e-6


def codebooks():
    rng = np.random.default_rng(0)
    e0 = rng.standard_normal((64, 512)).astype(np.float32)
    dead = e0.copy()
    dead[:, 60:] *= 1.0e5                      # collapsed regime: most codes are dead and huge
    small = (e0 * 1e-2).astype(np.float32)
    return {"randn": e0, "dead": dead, "small": small, "k256": e0[:, :256].copy()}


@pytest.mark.parametrize("cb", ["randn", "dead", "small", "k256"])
def test_tc_scores_are_certified_lower_bounds(cb):
    embed = codebooks()[cb]
    K = embed.shape[1]
    rng = np.random.default_rng(1)
    n = 1000                                   # ragged: 7 full tiles + 104 rows
    scale = 1e-2 if cb == "small" else 1.0
    x = (scale * rng.standard_normal((n, 64))).astype(np.float32)
    x[:64] = embed[:, rng.integers(0, min(K, 60), 64)].T          # rows equal to live codes
    scores, ind, flagged = tc_scores(x, embed)
    assert not np.isnan(scores).any(), "tensor-core scores were not written for every (row, code)"
    split, cA, cB = bound_constants()
    d64 = distances_f64(x, embed)
    xx = (x.astype(np.float64) ** 2).sum(1, keepdims=True)
    ee = (embed.astype(np.float64) ** 2).sum(0, keepdims=True)
    off = xx * (1.0 + 2.0 ** -9)
    target = d64 - xx + off                    # what the accumulator estimates before the bound is subtracted
    ebound = cA * np.sqrt(xx) * np.sqrt(ee) + cB * (ee + xx)
    tiny = 1e-6 * (xx + ee)
    err = scores.astype(np.float64) - target
    # lower bound, and not looser than twice the bound (+ bf16 upward roundings)
    assert (err <= tiny).all(), f"score above the true distance by {err.max():.3e} (bound violated)"
    assert (err >= -(2.2 * ebound + tiny)).all(), "score is looser than the documented bound"
    # how tight the filter really is (informational)
    raw = scores.astype(np.float64) + (cA * np.sqrt(xx) * np.sqrt(ee) + cB * ee) - target
    rel = np.abs(raw) / (np.sqrt(xx) * np.sqrt(ee) + 1e-30)
    print(f"[tc:{cb}] split={split} flagged={flagged}/{n} max|filter err|/(|x||e|)={rel.max():.3e} (budget cA/2={cA / 2:.3e})")
    # indices after the exact fix-up: exact-arithmetic arg-min
    o = QuantizeOracle(64, K, embed=embed)
    o.training = False
    _, _, io = o.forward(x)
    _, nbad, _ = tie_tolerant_index_mismatches(x, embed, ind, io)
    assert nbad == 0
    assert (ind >= 0).all() and (ind < K).all()
    assert flagged < n // 2 or split == 1


def test_tc_and_simt_engines_agree_on_outputs():
    torch.manual_seed(0)
    D, K, N = 64, 512, 128 * 37 + 5
    a = vq.Quantize(D, K, engine="tcgen05").to(DEV).train()
    b = vq.Quantize(D, K, engine="simt").to(DEV).train()
    b.load_state_dict(a.state_dict())
    for s in range(3):
        x = torch.randn(N, D, device=DEV, generator=torch.Generator(device=DEV).manual_seed(10 + s))
        qa, da, ia = a(x)
        qb, db, ib = b(x)
        ndiff, nbad, _ = tie_tolerant_index_mismatches(x.cpu().numpy(), b.embed.cpu().numpy(), ia.cpu().numpy(), ib.cpu().numpy())
        assert nbad == 0
        if int((ia != ib).sum()) == 0:
            assert torch.equal(qa, qb)
            assert abs(float(da) - float(db)) <= 1e-6 * float(db)
            assert rel_err(a.cluster_size.cpu().numpy(), b.cluster_size.cpu().numpy()) <= 1e-6
            assert col_rel_err(a.embed_avg.cpu().numpy(), b.embed_avg.cpu().numpy()) <= REL_TOL
        b.load_state_dict(a.state_dict())


def test_tc_exact_ties_and_duplicates_resolve_to_lowest_index():
    rng = np.random.default_rng(5)
    embed = rng.standard_normal((64, 512)).astype(np.float32)
    embed[:, 300] = embed[:, 7]
    embed[:, 450] = embed[:, 7]
    x = rng.standard_normal((256, 64)).astype(np.float32)
    x[:32] = embed[:, 7]
    x[32:64] = 0.5 * (embed[:, 20] + embed[:, 21])
    q = vq.Quantize(64, 512, engine="tcgen05").to(DEV).eval()
    q.embed.data.copy_(torch.from_numpy(embed))
    _, _, ind = q(torch.from_numpy(x).to(DEV))
    ind = ind.cpu().numpy()
    assert (ind[:32] == 7).all()
    o = QuantizeOracle(64, 512, embed=embed)
    o.training = False
    _, _, io = o.forward(x)
    _, nbad, _ = tie_tolerant_index_mismatches(x, embed, ind, io)
    assert nbad == 0


def test_tc_engine_rejects_uncovered_shapes_loudly():
    q = vq.Quantize(32, 40, engine="tcgen05").to(DEV).eval()
    with pytest.raises(RuntimeError, match="unsupported"):
        q(torch.randn(10, 32, device=DEV))


def test_tc_nonfinite_rows_do_not_poison_neighbours():
    torch.manual_seed(3)
    q = vq.Quantize(64, 512, engine="tcgen05").to(DEV).eval()
    x = torch.randn(300, 64, device=DEV)
    x[17, 3] = float("inf")
    x[130, 0] = float("nan")
    _, _, ind = q(x)
    ref = vq.Quantize(64, 512, engine="simt").to(DEV).eval()
    ref.load_state_dict(q.state_dict())
    _, _, ind_ref = ref(x)
    keep = torch.ones(300, dtype=torch.bool, device=DEV)
    keep[17] = keep[130] = False
    assert torch.equal(ind[keep], ind_ref[keep])
    assert int(ind.min()) >= 0 and int(ind.max()) < 512
