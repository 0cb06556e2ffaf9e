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


FILTERS = {"tcgen05": (3.0517578125e-5, 4.0e-6), "tcgen05_bf16": (7.9e-3, 4.0e-6), "tcgen05_tf32": (2.96e-3, 4.0e-6)}   # engine -> (cA, cB)


def tc_scores(x, embed, engine="tcgen05"):
    """Run vqb200_debug_tc_scores_ex; returns (scores [N,K], embed_ind [N], flagged_count)."""
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
    _native.check(lib.vqb200_debug_tc_scores_ex(_native.ptr(xd), n, d, k, _native.ptr(image), _native.ptr(ind),
                                                _native.ptr(scores), _native.ptr(flagged), _native.ptr(scratch),
                                                _native.ENGINES[engine], st), "scores")
    torch.cuda.synchronize()
    return scores.cpu().numpy(), ind.cpu().numpy(), int(flagged.item())


def codebooks():
    rng = np.random.default_rng(0)
    e0 = rng.standard_normal((64, 512)).astype(np.float32)
    dead = e0.copy()
    dead[:, 60:] *= 1.0e5                      # collapsed regime: most codes are dead and huge
    small = (e0 * 1e-2).astype(np.float32)
    return {"randn": e0, "dead": dead, "small": small, "k256": e0[:, :256].copy()}


@pytest.mark.parametrize("engine", ["tcgen05", "tcgen05_bf16", "tcgen05_tf32"])
@pytest.mark.parametrize("cb", ["randn", "dead", "small", "k256"])
def test_tc_scores_are_certified_lower_bounds(cb, engine):
    embed = codebooks()[cb]
    K = embed.shape[1]
    rng = np.random.default_rng(1)
    n = 1000                                   # ragged: 7 full tiles + 104 rows
    scale = 1e-2 if cb == "small" else 1.0
    x = (scale * rng.standard_normal((n, 64))).astype(np.float32)
    x[:64] = embed[:, rng.integers(0, min(K, 60), 64)].T          # rows equal to live codes
    scores, ind, flagged = tc_scores(x, embed, engine)
    assert not np.isnan(scores).any(), "tensor-core scores were not written for every (row, code)"
    cA, cB = FILTERS[engine]
    split = engine
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
    assert flagged < n // 2 or engine != "tcgen05"


@pytest.mark.parametrize("engine", ["tcgen05", "tcgen05_tf32"])
def test_tc_and_simt_engines_agree_on_outputs(engine):
    torch.manual_seed(0)
    D, K, N = 64, 512, 128 * 37 + 5
    a = vq.Quantize(D, K, engine=engine).to(DEV).train()
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


@pytest.mark.parametrize("engine", ["tcgen05_bf16", "tcgen05_tf32"])
def test_tc_plain_bf16_filter_is_exact_after_rescore(engine):
    """The cheap filters (plain bf16, tf32) certify fewer rows on N(0,1) inputs; the exact re-score must make the result
    identical to the SIMT engine anyway (heavy use of the flagged-row path)."""
    torch.manual_seed(21)
    D, K, N = 64, 512, 128 * 150 + 9
    a = vq.Quantize(D, K, engine=engine).to(DEV).train()
    b = vq.Quantize(D, K, engine="simt").to(DEV).train()
    b.load_state_dict(a.state_dict())
    x = torch.randn(N, D, device=DEV, generator=torch.Generator(device=DEV).manual_seed(6))
    qa, da, ia = a(x)
    qb, db, ib = b(x)
    assert int((ia != ib).sum()) == 0
    assert torch.equal(qa, qb)
    assert abs(float(da) - float(db)) <= 1e-6 * float(db)
    assert col_rel_err(a.embed_avg.cpu().numpy(), b.embed_avg.cpu().numpy()) <= REL_TOL


def test_auto_engine_adapts_filter_precision_without_changing_results():
    import copy
    torch.manual_seed(22)
    D, K, N = 64, 512, 128 * 150
    q = vq.Quantize(D, K).to(DEV).eval()
    ref = vq.Quantize(D, K, engine="simt").to(DEV).eval()
    ref.load_state_dict(q.state_dict())
    x = torch.randn(N, D, device=DEV)                      # many near-ties -> the policy must leave the bf16 filter
    modes = []
    for _ in range(4):
        _, _, ind = q(x)
        torch.cuda.synchronize()
        modes.append(q._filter["mode"])
        assert torch.equal(ind, ref(x)[2])
    assert "split" in modes
    q2 = copy.deepcopy(q)                                  # pending CUDA event / workspaces must not break deepcopy
    assert torch.equal(q2(x)[2], ref(x)[2])


@pytest.mark.parametrize("engine", ["tcgen05", "tcgen05_tf32"])
@pytest.mark.parametrize("K", [512, 256])
def test_tc_many_trips_per_cta(K, engine):
    """More tiles than 3x the SM count: every persistent CTA makes several trips, so the stage / TMEM-buffer
    phase logic (and the CTA-pair tail, where the second tile of the last pair lies past the end) is exercised."""
    torch.manual_seed(11)
    D, N = 64, 128 * 148 * 3 + 128 + 77
    a = vq.Quantize(D, K, engine=engine).to(DEV).eval()
    b = vq.Quantize(D, K, engine="simt").to(DEV).eval()
    b.load_state_dict(a.state_dict())
    x = torch.randn(N, D, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
    qa, da, ia = a(x)
    qb, db, ib = b(x)
    _, nbad, _ = tie_tolerant_index_mismatches(x.cpu().numpy(), b.embed.cpu().numpy(), ia.cpu().numpy(), ib.cpu().numpy())
    assert nbad == 0
    if int((ia != ib).sum()) == 0:
        assert torch.equal(qa, qb)
        assert abs(float(da) - float(db)) <= 1e-6 * float(db)


@pytest.mark.parametrize("engine", ["tcgen05", "tcgen05_bf16", "tcgen05_tf32"])
@pytest.mark.parametrize("live_codes", [1, 3, 40])
def test_tc_fused_statistics_with_skewed_codes(engine, live_codes):
    """Heavily skewed assignments (collapsed codebook: a handful of live codes): the per-code sums and
    counts behind the tensor-core engine must equal the float64 segmented sums."""
    torch.manual_seed(31)
    D, K, N = 64, 512, 128 * 301 + 17
    q = vq.Quantize(D, K, engine=engine).to(DEV).train()
    embed0 = q.embed.clone()
    live = torch.randperm(K, device=DEV)[:live_codes]
    pick = live[torch.randint(0, live_codes, (N,), device=DEV)]
    x = (embed0.t()[pick] + 0.05 * torch.randn(N, D, device=DEV)).contiguous()
    _, _, ind = q(x)
    assert torch.equal(ind, pick)
    # float64 segmented sums of what the kernel must have accumulated
    sums = torch.zeros(K, D, dtype=torch.float64, device=DEV).index_add_(0, ind, x.double())
    counts = torch.bincount(ind, minlength=K).double()
    decay = 0.99
    want_cs = counts * (1 - decay)
    want_avg = embed0.double() * decay + sums.t() * (1 - decay)
    assert torch.allclose(q.cluster_size.double(), want_cs, rtol=1e-6, atol=0)
    err = (q.embed_avg.double() - want_avg).abs().max(0).values / want_avg.abs().max(0).values
    assert float(err.max()) <= REL_TOL


@pytest.mark.parametrize("shape", [(3, 64, 5, 7), (4, 64, 32, 32)])
def test_tc_engine_on_nchw_physical_input(shape):
    """VQVAE.encode passes conv_output.permute(0, 2, 3, 1) (vqvae.py:227,235): rows with unit row stride.  The module
    re-packs such inputs to dense rows for the tcgen05 engine and writes `quantize` back with the input's strides."""
    torch.manual_seed(41)
    B, D, H, W = shape
    a = vq.Quantize(D, 512).to(DEV).train()
    b = vq.Quantize(D, 512, engine="simt").to(DEV).train()
    b.load_state_dict(a.state_dict())
    x = torch.randn(B, D, H, W, device=DEV).permute(0, 2, 3, 1)
    assert not x.is_contiguous()
    qa, da, ia = a(x)
    qb, db, ib = b(x)
    assert qa.stride() == x.stride() and qa.shape == x.shape          # vqvae.py:73 keeps the input's strides
    assert torch.equal(ia, ib)
    assert torch.equal(qa, qb)
    assert abs(float(da) - float(db)) <= 1e-6 * float(db)
    assert col_rel_err(a.embed_avg.cpu().numpy(), b.embed_avg.cpu().numpy()) <= REL_TOL
    assert torch.allclose(a.cluster_size, b.cluster_size, rtol=1e-6, atol=0)
    assert _native.load().vqb200_launch_count() > 0


@pytest.mark.parametrize("engine", ["tcgen05", "tcgen05_bf16", "tcgen05_tf32"])
@pytest.mark.parametrize("K", [1024, 2048])
def test_tc_sliced_codebook_matches_simt(K, engine):
    """Codebooks larger than the resident operand image (BASELINE cfg-5 sweep, D = 64): one tensor-core launch per
    512-code slice, the running (best, runner-up, winner) carried per row; only the last slice certifies and writes."""
    torch.manual_seed(51)
    D, N = 64, 128 * 37 + 45
    a = vq.Quantize(D, K, engine=engine).to(DEV).train()
    b = vq.Quantize(D, K, engine="simt").to(DEV).train()
    b.load_state_dict(a.state_dict())
    embed0 = a.embed.clone()
    pick = torch.randint(0, K, (N,), device=DEV)
    x = torch.cat([embed0.t()[pick[: N // 2]] + 0.2 * torch.randn(N // 2, D, device=DEV),
                   torch.randn(N - N // 2, D, device=DEV)]).contiguous()       # half clustered, half N(0,1) (many near-ties)
    for step in range(2):                      # second step: codebook with dead ~1e5-magnitude codes
        embed_before = b.embed.cpu().numpy().copy()
        qa, da, ia = a(x)
        qb, db, ib = b(x)
        _, nbad, _ = tie_tolerant_index_mismatches(x.cpu().numpy(), embed_before, ia.cpu().numpy(), ib.cpu().numpy())
        assert nbad == 0
        if int((ia != ib).sum()) == 0:
            assert torch.equal(qa, qb)
            assert abs(float(da) - float(db)) <= 1e-5 * abs(float(db))
            assert col_rel_err(a.embed_avg.cpu().numpy(), b.embed_avg.cpu().numpy()) <= REL_TOL
            assert torch.allclose(a.cluster_size, b.cluster_size, rtol=1e-5, atol=1e-7)
        b.load_state_dict(a.state_dict())


@pytest.mark.parametrize("engine", ["tcgen05", "tcgen05_bf16"])
@pytest.mark.parametrize("K", [256, 512, 1024])
@pytest.mark.parametrize("shape,train", [((4, 64, 32, 32), True), ((1, 64, 128, 128), True), ((37, 64, 16, 8), False)])
def test_tc_nchw_in_place_variants(shape, train, K, engine):
    """The NCHW-physical kernel variant (3-D TMA tensor-map tile loads, x^T stage, row-per-thread converters, row-coalesced
    output) over both filters (incl. the CTA-pair split kernel), the sliced codebook, training (dense copy for the
    statistics) and eval, many trips per CTA and a single image."""
    torch.manual_seed(61)
    B, D, H, W = shape
    a = vq.Quantize(D, K, engine=engine).to(DEV).train(train)
    b = vq.Quantize(D, K, engine="simt").to(DEV).train(train)
    b.load_state_dict(a.state_dict())
    embed0 = a.embed.clone()
    n = B * H * W
    pick = torch.randint(0, K, (n,), device=DEV)
    dense = embed0.t()[pick] + 0.25 * torch.randn(n, D, device=DEV)
    x = dense.reshape(B, H, W, D).permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)     # NCHW-physical
    lib = _native.load()
    from vq_vae_2_pytorch_b200 import row_layout
    lay = row_layout(x)
    assert lib.vqb200_tc_supported(_native.ptr(x), lay[0], D, K, lay[1], lay[2], lay[3], lay[4]) == 1   # consumed in place
    for step in range(2 if train else 1):
        embed_before = b.embed.cpu().numpy().copy()
        qa, da, ia = a(x)
        qb, db, ib = b(x)
        assert qa.stride() == x.stride()
        _, nbad, _ = tie_tolerant_index_mismatches(x.cpu().numpy(), embed_before, ia.cpu().numpy(), ib.cpu().numpy())
        assert nbad == 0
        if int((ia != ib).sum()) == 0:
            assert torch.equal(qa, qb)
            assert abs(float(da) - float(db)) <= 1e-5 * abs(float(db))
            if train:
                assert col_rel_err(a.embed_avg.cpu().numpy(), b.embed_avg.cpu().numpy()) <= REL_TOL
                assert torch.allclose(a.cluster_size, b.cluster_size, rtol=1e-5, atol=1e-7)
        b.load_state_dict(a.state_dict())


# ---------------------------------------------------------------------------------------------------------------------
# The certificate's accumulation constant cB is measured, not proven: the tensor core's internal
# accumulation is not IEEE fp32.  This sweep isolates it: operands that every filter represents EXACTLY (integers / 4 with
# |v| <= 2, times a power-of-two scale: 4 significant bits, so bf16, tf32 and the bias / offset splits are all exact) leave
# the accumulation as the only error source.  Magnitudes 1e-6 .. 1e6 of both x and e (and mixed), N = 1e5 rows per point,
# dims 64 / 128 / 256; the worst |score error| / (||x||^2 + ||e||^2) is printed, written to gpurun_out/, and must stay below
# cB / 2 (the certificate budgets cB for it; a 2x margin is asserted on every GPU run).
# ---------------------------------------------------------------------------------------------------------------------
SWEEP = [("tcgen05", 64, 512), ("tcgen05_bf16", 64, 512), ("tcgen05_tf32", 64, 512), ("tcgen05_tf32", 64, 256),
         ("tcgen05", 128, 512), ("tcgen05", 256, 256)]


@pytest.mark.parametrize("engine,D,K", SWEEP)
def test_certificate_accumulation_constant_sweep(engine, D, K):
    import json
    import os
    rng = np.random.default_rng(11)
    cB = 4.0e-6 * (D // 64)
    worst = {}
    for sx, se in [(0, 0), (-20, -20), (20, 20), (-20, 0), (0, -20), (0, 17), (10, -10), (-10, 10), (17, 17)]:
        n = (100000 if (sx, se) == (0, 0) else 16384) // 128 * 128 + 96      # >= 1e5 rows at unit scale, 16 k at the others
        x = (rng.integers(-8, 9, size=(n, D)) / 4.0 * 2.0 ** sx).astype(np.float32)
        e = (rng.integers(-8, 9, size=(D, K)) / 4.0 * 2.0 ** se).astype(np.float32)
        e[:, K // 2:] *= np.float32(2.0 ** 3)             # two magnitude classes inside one accumulator unit
        scores, ind, flagged = tc_scores(x, e, engine)
        xx = (x.astype(np.float64) ** 2).sum(1, keepdims=True)
        ee = (e.astype(np.float64) ** 2).sum(0, keepdims=True)
        cA = FILTERS[engine][0] if D == 64 else 7.9e-3
        # what the contraction computes when nothing but the accumulation rounds: bias + offset - 2 x.e - bound terms
        exact = ee - 2.0 * (x.astype(np.float64) @ e.astype(np.float64)) + xx * (1.0 + 2.0 ** -9)
        resid = scores.astype(np.float64) - exact          # = -(bound terms, rounded up) + accumulation error
        bound_lo = -(cA * 1.03125 * np.sqrt(xx) * np.sqrt(ee) + cB * 1.02 * ee)      # BOUND_UP covers the upward roundings
        acc_err_up = resid                                  # the score may exceed `exact` only through accumulation error
        acc_err_dn = bound_lo - resid                       # or fall below the nominal bound terms only through it
        ratio = np.maximum(acc_err_up, acc_err_dn) / (xx + ee)
        worst[f"x2^{sx} e2^{se}"] = float(ratio.max())
        assert ratio.max() <= cB / 2, f"{engine} D={D} K={K} scale (2^{sx}, 2^{se}): accumulation error {ratio.max():.3e} (xx+ee) > cB/2"
        assert (ind >= 0).all() and (ind < K).all()
    print(f"[cB sweep {engine} D={D} K={K}] worst accumulation error / (xx+ee) = {max(worst.values()):.3e} (cB = {cB:.1e})")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, f"cb_sweep_{engine}_D{D}_K{K}.json"), "w") as f:
            json.dump({"engine": engine, "dim": D, "n_embed": K, "rows_unit_scale": 100000 // 128 * 128 + 96, "cB": cB,
                       "worst_ratio_by_scale": worst}, f, indent=1)
    except OSError:
        pass
