"""GPU parity against THE REFERENCE ITSELF: the unmodified `Quantize` / `VQVAE` / `VQVAE_Deep` of
alehdaghi/vq-vae-2-pytorch (staged byte for byte in oracle/_ref by tools/fetch_ref.py, digests pinned in
oracle/ref_manifest.json) executed on the same B200 with TF32 off, against the CUDA path behind the drop-in module --
at BASELINE.json's full sizes (north_star: "Correctness is checked against the reference's own PyTorch Quantize on
identical synthetic inputs"):

  cfg-2  bottom quantizer [128,64,64,64], D=64, K=512, train, dense and NCHW-physical, randn -> collapsed -> clustered
  cfg-3  top [256,32,32,64] then bottom [256,64,64,64] on one module pair, two optimiser steps' worth
  cfg-4  eval, B=1024, top + bottom (4.19 M rows), argmin-only and full eval forward
  cfg-5  corners (64,8192) (128,8192) (256,8192) and interior points, N = 524 288, row-chunked reference
  the unmodified VQVAE / VQVAE_Deep with only the `Quantize` class swapped (vqvae.py:185,190; vqvae_deep.py:252,257),
  through encode()/quantize(), forward+backward, and the extract_code.py:14-33 loop

Tolerances and the near-tie rule: tests/ref_harness.py.
"""
import numpy as np
import pytest
import torch

import vq_vae_2_pytorch_b200 as vq
import ref_harness as H
from oracle import reference_module

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ref():
    try:
        mod = reference_module.load("vqvae")
    except reference_module.ReferenceUnavailable as exc:
        pytest.skip(f"reference not staged: {exc}")
    H.fp32_reference_backends()
    yield mod
    H.dump_report()


def pair(ref, D, K, seed=0, engine="auto", train=True):
    torch.manual_seed(seed)
    r = ref.Quantize(D, K).to(DEV).train(train)
    o = vq.Quantize(D, K, engine=engine).to(DEV).train(train)
    o.load_state_dict(r.state_dict(), strict=True)          # reference checkpoints load strictly
    return r, o


def free_memory():
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


# --------------------------------------------------------------------------------------------- cfg-2
@pytest.mark.parametrize("permuted", [False, True], ids=["dense", "nchw"])
def test_cfg2_full_size_multistep_vs_reference(ref, permuted):
    """randn (reference-init regime, ~489 codes hit) -> randn on the collapsed codebook (dead codes ~1e5) -> clustered
    around the live codes -> clustered again; full size, free-running EMA on the reference's trajectory."""
    B, Hh, W, D, K = 128, 64, 64, 64, 512
    r, o = pair(ref, D, K)
    kinds = ["randn", "randn", "clustered", "clustered"]
    for step, kind in enumerate(kinds):
        x = H.make_inputs(kind, (B, Hh, W, D), r.embed.detach(), 1234 + 1000 * step, DEV, permuted)
        e = H.compare_step(f"cfg2-{'nchw' if permuted else 'dense'}-step{step}-{kind}", r, o, x)
        assert e["index_differ_near_tie"] <= 4
    assert float(r.embed.abs().max()) > 1e4                  # the dead-code blow-up really happened
    free_memory()


def test_cfg2_filter_engines_vs_reference(ref):
    """Both tensor-core filters pinned explicitly (the precision policy of engine='auto' may pick either) + the SIMT engine."""
    D, K = 64, 512
    for engine in ("tcgen05", "tcgen05_bf16", "tcgen05_tf32", "simt"):
        r, o = pair(ref, D, K, seed=3, engine=engine)
        for step, kind in enumerate(["randn", "clustered"]):
            x = H.make_inputs(kind, (32, 64, 64, D), r.embed.detach(), 99 + step, DEV)
            H.compare_step(f"cfg2-quarter-{engine}-step{step}-{kind}", r, o, x)
    free_memory()


# --------------------------------------------------------------------------------------------- cfg-3
def test_cfg3_top_then_bottom_training_sequence_vs_reference(ref):
    """One optimiser step = top forward then bottom forward (vqvae.py:227-237), B = 256, NCHW-physical inputs as the
    model passes them; two steps, each module sees its own previous update."""
    D, K, B = 64, 512, 256
    rt, ot = pair(ref, D, K, seed=1)
    rb, ob = pair(ref, D, K, seed=2)
    for step in range(2):
        kind = "randn" if step == 0 else "clustered"
        xt = H.make_inputs(kind, (B, 32, 32, D), rt.embed.detach(), 500 + step, DEV, permuted=True)
        H.compare_step(f"cfg3-top-step{step}-{kind}", rt, ot, xt)
        xb = H.make_inputs(kind, (B, 64, 64, D), rb.embed.detach(), 600 + step, DEV, permuted=True)
        H.compare_step(f"cfg3-bottom-step{step}-{kind}", rb, ob, xb)
    free_memory()


# --------------------------------------------------------------------------------------------- cfg-4
@pytest.mark.parametrize("kind", ["clustered", "randn"])
def test_cfg4_inference_b1024_vs_reference(ref, kind):
    """extract_code.py-style inference (eval, no EMA), B = 1024: top 1 048 576 rows + bottom 4 194 304 rows; the reference
    runs in 524 288-row chunks (rows are independent in eval mode).  Also Quantize.assign (indices only)."""
    D, K = 64, 512
    r, o = pair(ref, D, K, seed=4, train=False)
    for name, shape in (("top", (1024, 32, 32, D)), ("bottom", (1024, 64, 64, D))):
        x = H.make_inputs(kind, shape, r.embed.detach(), 4242, DEV, permuted=True)
        e = H.compare_step(f"cfg4-{name}-{kind}", r, o, x, ref_chunk=524288)
        ind = o.assign(x)
        with torch.no_grad():
            assert torch.equal(ind, o(x)[2])
        assert e["index_differ_near_tie"] <= 8
        del x, ind
        free_memory()


# --------------------------------------------------------------------------------------------- cfg-5
@pytest.mark.parametrize("D,K", [(64, 8192), (128, 8192), (256, 8192), (64, 2048), (128, 1024), (256, 512), (256, 1024), (64, 4096)])
def test_cfg5_sweep_points_vs_chunked_reference(ref, D, K):
    """Codebook sweep at N = 524 288 (B = 128, 64x64): train step 0 on half clustered / half N(0,1) rows (near-ties and the
    exact fix-up), train step 1 on the collapsed codebook, against the row-chunked reference.  Covers the K/512-slice
    carry (`partial`) at 16 slices and the streamed-operand path of the wide engine."""
    N = 128 * 64 * 64
    r, o = pair(ref, D, K, seed=5)
    chunk = 65536 if K >= 4096 else 131072
    for step in range(2):
        g = torch.Generator(device=DEV).manual_seed(7000 + step)
        xa = H.make_inputs("clustered", (N // 2, D), r.embed.detach(), 7100 + step, DEV)
        xb = torch.randn(N - N // 2, D, device=DEV, generator=g)
        x = torch.cat([xa, xb]).contiguous()
        del xa, xb
        e = H.compare_step(f"cfg5-D{D}-K{K}-step{step}", r, o, x, ref_chunk=chunk)
        assert e["index_differ_near_tie"] <= 8
        del x
        free_memory()


# --------------------------------------------------------------------------------------------- drop-in: unmodified VQVAE
def build_models(refmod, cls_name, **kw):
    """The reference model twice from the same seed: as shipped, and with the module-level `Quantize` swapped for the
    drop-in (`vqvae.Quantize = vq_vae_2_pytorch_b200.Quantize`).  Nothing else is touched."""
    cls = getattr(refmod, cls_name)
    torch.manual_seed(0)
    ref_model = cls(**kw).to(DEV)
    orig = refmod.Quantize
    refmod.Quantize = vq.Quantize
    try:
        torch.manual_seed(0)
        our_model = cls(**kw).to(DEV)
    finally:
        refmod.Quantize = orig
    assert isinstance(our_model.quantize_t, vq.Quantize) and isinstance(our_model.quantize_b, vq.Quantize)
    assert isinstance(ref_model.quantize_t, orig)
    missing = our_model.load_state_dict(ref_model.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return ref_model, our_model


def test_unmodified_vqvae_with_swapped_quantize(ref):
    """VQVAE.encode (vqvae.py:223-240) and a full forward+backward training step, reference vs class-swapped model."""
    torch.backends.cudnn.deterministic = True
    ref_model, our_model = build_models(ref, "VQVAE")
    ref_model.train(); our_model.train()
    opt_r = torch.optim.Adam(ref_model.parameters(), lr=3e-4)
    opt_o = torch.optim.Adam(our_model.parameters(), lr=3e-4)
    for step in range(3):
        img = torch.randn(8, 3, 256, 256, device=DEV, generator=torch.Generator(device=DEV).manual_seed(40 + step))
        outs = []
        for model, opt in ((ref_model, opt_r), (our_model, opt_o)):
            opt.zero_grad()
            quant_t, quant_b, diff, id_t, id_b = model.encode(img)                 # vqvae.py:223-240
            upsample_t = model.upsample_t(quant_t)
            dec = model.decode(torch.cat([upsample_t, quant_b], 1))
            loss = (dec - img).pow(2).mean() + 0.25 * diff.mean()                   # train_vqvae.py: recon + 0.25 * latent
            loss.backward()
            outs.append((quant_t.detach(), quant_b.detach(), diff.detach(), id_t, id_b, loss.detach(),
                         model.enc_b.blocks[0].weight.grad.detach().clone(), model.quantize_conv_b.weight.grad.detach().clone()))
            opt.step()
        (rt, rb, rd, rit, rib, rl, rg0, rg1), (ot, ob, od, oit, oib, ol, og0, og1) = outs
        assert rt.shape == (8, 64, 32, 32) and rb.shape == (8, 64, 64, 64) and rit.shape == (8, 32, 32) and rib.shape == (8, 64, 64)
        assert ot.stride() == rt.stride() and ob.stride() == rb.stride()
        assert torch.equal(oit, rit), f"step {step}: top ids differ in {int((oit != rit).sum())} places"
        assert torch.equal(oib, rib), f"step {step}: bottom ids differ in {int((oib != rib).sum())} places"
        assert torch.allclose(ot, rt, rtol=1e-5, atol=1e-6) and torch.allclose(ob, rb, rtol=1e-5, atol=1e-6)
        assert torch.allclose(od, rd, rtol=1e-5, atol=0) and torch.allclose(ol, rl, rtol=1e-5, atol=0)
        for a, b in ((og0, rg0), (og1, rg1)):
            assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max()) + 1e-12
        for name in ("quantize_t", "quantize_b"):
            qo, qr = getattr(our_model, name), getattr(ref_model, name)
            assert torch.allclose(qo.cluster_size, qr.cluster_size, rtol=1e-5, atol=1e-7)
            assert H.scaled_err(qo.embed_avg, qr.embed_avg, qr.embed_avg.abs().amax(0, keepdim=True).double().clamp_min(1e-30)) <= 1e-5
            H.sync_state(qo, qr)
        # keep the two optimisers' parameters bit-identical so that later steps compare the quantizers, not Adam noise
        our_model.load_state_dict(ref_model.state_dict(), strict=True)
    # extract_code.py:14-33: eval, encode, ids to the host
    ref_model.eval(); our_model.eval()
    with torch.no_grad():
        img = torch.randn(16, 3, 256, 256, device=DEV, generator=torch.Generator(device=DEV).manual_seed(77))
        _, _, _, rit, rib = ref_model.encode(img)
        _, _, _, oit, oib = our_model.encode(img)
        assert np.array_equal(oit.detach().cpu().numpy(), rit.detach().cpu().numpy())
        assert np.array_equal(oib.detach().cpu().numpy(), rib.detach().cpu().numpy())
        # decode side (vqvae.py:251-255 up to the fork's broken decode() arity): embed_code on both quantizers
        assert torch.equal(our_model.quantize_t.embed_code(oit), ref_model.quantize_t.embed_code(rit))
        assert torch.equal(our_model.quantize_b.embed_code(oib), ref_model.quantize_b.embed_code(rib))
    torch.backends.cudnn.deterministic = False
    free_memory()


def test_unmodified_vqvae_deep_quantize_call_sites(ref):
    """vqvae_deep.py:288-299: D = 256 quantizers on permuted [B,256,36,18] / [B,256,18,9] views of 288x144 inputs."""
    try:
        deep = reference_module.load("vqvae_deep")
    except reference_module.ReferenceUnavailable as exc:
        pytest.skip(str(exc))
    ref_model, our_model = build_models(deep, "VQVAE_Deep")
    ref_model.train(); our_model.train()
    for step in range(2):
        img = torch.randn(8, 3, 288, 144, device=DEV, generator=torch.Generator(device=DEV).manual_seed(50 + step))
        with torch.no_grad():
            rt, rb, rd, rit, rib = ref_model.quantize(*ref_model.encode(img))
            ot, ob, od, oit, oib = our_model.quantize(*our_model.encode(img))
        assert rit.shape[1:] == (18, 9) and rib.shape[1:] == (36, 18) and rt.shape[1] == 256
        assert torch.equal(oit, rit) and torch.equal(oib, rib)
        assert ot.stride() == rt.stride() and ob.stride() == rb.stride()
        assert torch.allclose(ot, rt, rtol=1e-5, atol=1e-6) and torch.allclose(ob, rb, rtol=1e-5, atol=1e-6)
        assert torch.allclose(od, rd, rtol=1e-5, atol=0)
        for name in ("quantize_t", "quantize_b"):
            qo, qr = getattr(our_model, name), getattr(ref_model, name)
            assert torch.allclose(qo.cluster_size, qr.cluster_size, rtol=1e-5, atol=1e-7)
            assert H.scaled_err(qo.embed_avg, qr.embed_avg, qr.embed_avg.abs().amax(0, keepdim=True).double().clamp_min(1e-30)) <= 1e-5
            H.sync_state(qo, qr)
    free_memory()


def test_channels_last_model_feeds_the_dense_fast_path(ref):
    """INTEGRATION.md recipe: run the (unmodified, class-swapped) reference VQVAE in torch.channels_last.  The 1x1
    `quantize_conv_*` outputs are then NHWC-physical, so the `permute(0, 2, 3, 1)` views of vqvae.py:227,235 are CONTIGUOUS
    [B,H,W,D] rows -- the quantizer's dense fast path, no strided kernel variant -- and `quantize.permute(0, 3, 1, 2)` is a
    channels_last tensor the decoder convolutions consume natively.  Results must equal the NCHW run of the reference."""
    from vq_vae_2_pytorch_b200 import row_layout
    ref_model, our_model = build_models(ref, "VQVAE")
    ref_model.eval()
    our_model = our_model.to(memory_format=torch.channels_last).eval()
    seen = []
    hooks = [m.register_forward_pre_hook(lambda mod, args: seen.append((args[0].is_contiguous(), row_layout(args[0]))))
             for m in (our_model.quantize_t, our_model.quantize_b)]
    img = torch.randn(8, 3, 256, 256, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
    with torch.no_grad():
        rt, rb, rd, rit, rib = ref_model.encode(img)
        ot, ob, od, oit, oib = our_model.encode(img.contiguous(memory_format=torch.channels_last))
    for h in hooks:
        h.remove()
    assert len(seen) == 2 and all(c for c, _ in seen), "quantizer inputs are not contiguous under channels_last"
    for _, lay in seen:
        n, rpi, img_stride, row, col = lay
        assert (row, col) == (64, 1)                                          # dense rows: the fast path
    assert ot.is_contiguous(memory_format=torch.channels_last) and ob.is_contiguous(memory_format=torch.channels_last)
    # cuDNN may pick other algorithms for the other memory format: the latents agree to conv round-off, the ids wherever
    # the two runs' latents do not straddle a near-tie
    assert float((oit != rit).float().mean()) <= 1e-3 and float((oib != rib).float().mean()) <= 1e-3
    same_t = (oit == rit).unsqueeze(1).expand_as(rt)
    assert torch.allclose(ot[same_t], rt[same_t], rtol=1e-4, atol=1e-5)
    assert torch.allclose(od, rd, rtol=1e-3, atol=0)
    free_memory()
