"""GPU parity tests: the CUDA path (through the C ABI, via the drop-in module) against
 (1) the golden fixtures = outputs of the reference `Quantize` (tests/golden/make_golden.py), and
 (2) the CPU oracle (oracle/quantize_oracle.py) on seeded inputs, plus size-independent properties at
     BASELINE.json's full sizes.

Tolerances (BASELINE.json): embed_ind exact except fp32 near-ties (float64 distance gap < 1e-6
relative to the magnitude of the summed terms); quantize / diff / EMA buffers within 1e-5 relative.
EMA buffers are compared only when the indices agree (one differing index moves a whole vector
between two codebook columns).
"""
import ctypes as C

import numpy as np
import pytest
import torch

import vq_vae_2_pytorch_b200 as vq
from vq_vae_2_pytorch_b200 import _native
from helpers import REL_TOL, check_outputs_np, col_rel_err, elem_rel_err, golden_names, load_golden, rel_err
from oracle.quantize_oracle import QuantizeOracle, tie_tolerant_index_mismatches

pytestmark = pytest.mark.gpu
ENGINES = ["simt", "auto"]
DEV = "cuda:0"


def make_module(g_or_dims, engine, embed=None, cluster_size=None, embed_avg=None, decay=0.99, eps=1e-5):
    dim, n_embed = g_or_dims
    q = vq.Quantize(dim, n_embed, decay=decay, eps=eps, engine=engine).to(DEV)
    if embed is not None:
        q.embed.data.copy_(torch.from_numpy(embed))
        q.embed_avg.data.copy_(torch.from_numpy(embed if embed_avg is None else embed_avg))
    if cluster_size is not None:
        q.cluster_size.data.copy_(torch.from_numpy(cluster_size))
    return q


def as_strided_cuda(arr, strides):
    """Rebuild a tensor with the reference run's strides (element units) on the GPU."""
    t = torch.from_numpy(np.ascontiguousarray(arr))
    out = torch.empty_strided(t.shape, tuple(int(s) for s in strides), dtype=t.dtype, device=DEV)
    out.copy_(t)
    return out


def check_outputs(tag, x_np, state_before, outs, want, module, want_state, train, decay=0.99, eps=1e-5):
    """torch front of helpers.check_outputs_np (element-wise 1e-5, near-tie rule, touched EMA columns excluded)."""
    quant, diff, ind = outs
    assert ind.dtype == torch.int64 and ind.is_contiguous() and tuple(ind.shape) == x_np.shape[:-1]
    assert quant.dtype == torch.float32 and diff.dim() == 0 and diff.dtype == torch.float32
    got_state = (module.cluster_size.cpu().numpy(), module.embed_avg.cpu().numpy(), module.embed.cpu().numpy()) if train else None
    return check_outputs_np(tag, x_np, state_before, (quant.detach().cpu().numpy(), float(diff), ind.cpu().numpy()), want,
                            got_state, want_state, train, decay, eps)


def module_state(q):
    return {k: v.detach().cpu().numpy().copy() for k, v in q.state_dict().items()}


def load_state(q, cluster_size, embed_avg, embed):
    q.embed.data.copy_(torch.from_numpy(np.ascontiguousarray(embed)))
    q.cluster_size.data.copy_(torch.from_numpy(np.ascontiguousarray(cluster_size)))
    q.embed_avg.data.copy_(torch.from_numpy(np.ascontiguousarray(embed_avg)))


def check_step(name, s, x_np, state_before, outs, golden, module, check_buffers=True):
    want_state = (golden[f"cluster_size{s + 1}"], golden[f"embed_avg{s + 1}"], golden[f"embed{s + 1}"]) if check_buffers else None
    return check_outputs(f"{name} step {s}", x_np, state_before, outs,
                         (golden[f"quantize{s}"], golden[f"diff{s}"], golden[f"ind{s}"]), module, want_state, check_buffers,
                         float(golden["decay"]), float(golden["eps"]))


def oracle_step(tag, q, o, x_np, x_dev=None):
    """One forward of module and oracle from the same state, element-wise check, then the module follows the oracle's
    trajectory (so a tolerated near-tie never forks the sequence: no skips)."""
    before = {"embed": o.embed.copy(), "embed_avg": o.embed_avg.copy(), "cluster_size": o.cluster_size.copy()}
    qo, do, io = o.forward(x_np)
    outs = q(torch.from_numpy(x_np).to(DEV) if x_dev is None else x_dev)
    nd = check_outputs(tag, x_np, before, outs, (qo, do, io), q, (o.cluster_size, o.embed_avg, o.embed), q.training,
                       o.decay, o.eps)
    if q.training:
        load_state(q, o.cluster_size, o.embed_avg, o.embed)
    return nd, outs


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", golden_names())
def test_golden_fixture(name, engine):
    g = load_golden(name)
    dims = (int(g["dim"]), int(g["n_embed"]))
    q = make_module(dims, engine, g["embed0"], g["cluster_size0"], g["embed_avg0"], float(g["decay"]), float(g["eps"]))
    q.train(bool(g["train"]))
    for s in range(int(g["steps"])):
        x_np = g[f"x{s}"]
        x = as_strided_cuda(x_np, g[f"x{s}_strides"])
        want_grad = f"xgrad{s}" in g
        x.requires_grad_(want_grad)
        state_np = module_state(q)
        state_before = {k: v.clone() for k, v in q.state_dict().items()}
        quant, diff, ind = q(x)
        assert tuple(quant.stride()) == tuple(int(v) for v in g[f"quantize{s}_strides"])   # strides follow the input
        ndiff = check_step(name, s, np.ascontiguousarray(x_np), state_np, (quant.detach(), diff.detach(), ind), g, q,
                           check_buffers=bool(g["train"]))
        if not g["train"]:
            for k, v in q.state_dict().items():
                assert torch.equal(v, state_before[k]), "eval mode must not touch the buffers (vqvae.py:54)"
        if want_grad:
            gq = torch.from_numpy(g[f"gq{s}"]).to(DEV)
            (quant * gq).sum().add(diff * float(g[f"gd{s}"])).backward()
            keep_rows = (ind.cpu().numpy() == g[f"ind{s}"]).reshape(-1)
            got, want = x.grad.cpu().numpy().reshape(-1, dims[0])[keep_rows], g[f"xgrad{s}"].reshape(-1, dims[0])[keep_rows]
            assert elem_rel_err(got, want, floor=1e-6 * float(np.abs(want).max())) <= REL_TOL
        # follow the reference trajectory so later steps stay comparable even after a tolerated near-tie
        q.embed.data.copy_(torch.from_numpy(g[f"embed{s + 1}"]))
        q.cluster_size.data.copy_(torch.from_numpy(g[f"cluster_size{s + 1}"]))
        q.embed_avg.data.copy_(torch.from_numpy(g[f"embed_avg{s + 1}"]))


@pytest.mark.parametrize("engine", ENGINES)
def test_free_running_ema_sequence_matches_oracle(engine):
    """5 training steps without re-synchronising: the collapsed-codebook regime (dead codes ~1e5)."""
    D, K = 64, 512
    torch.manual_seed(0)
    q = vq.Quantize(D, K, engine=engine).to(DEV).train()
    o = QuantizeOracle(D, K, embed=q.embed.cpu().numpy())
    for s in range(5):
        x = torch.randn(4, 32, 32, D, generator=torch.Generator().manual_seed(1234 + 1000 * s))
        oracle_step(f"free-running step {s}", q, o, x.numpy())
    assert float(q.embed.abs().max()) > 1e4          # the dead-code blow-up really happened


@pytest.mark.parametrize("engine", ENGINES)
def test_clustered_and_scaled_inputs_vs_oracle(engine):
    D, K = 64, 512
    rng = np.random.default_rng(3)
    embed = rng.standard_normal((D, K)).astype(np.float32)
    for scale, noise in ((1.0, 0.1), (30.0, 2.0), (1e-3, 1e-4)):
        e = (embed * scale).astype(np.float32)
        pick = rng.integers(0, K, size=3000)
        x = (e[:, pick].T + noise * rng.standard_normal((3000, D))).astype(np.float32)
        q = make_module((D, K), engine, e).train()
        o = QuantizeOracle(D, K, embed=e)
        oracle_step(f"clustered scale {scale}", q, o, x)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("shape,D,K", [((1, 64), 64, 512), ((127, 64), 64, 512), ((129, 64), 64, 512),
                                       ((3, 5, 32), 32, 40), ((1000, 256), 256, 1024), ((77, 128), 128, 2048),
                                       ((300, 48), 48, 100), ((64, 8), 8, 3), ((50, 3), 3, 5), ((33, 7), 7, 9)])
def test_ragged_shapes_vs_oracle(shape, D, K, engine):
    rng = np.random.default_rng(hash((shape, D, K)) % (2 ** 31))
    embed = rng.standard_normal((D, K)).astype(np.float32)
    x = rng.standard_normal(shape).astype(np.float32)
    q = make_module((D, K), engine, embed).train()
    o = QuantizeOracle(D, K, embed=embed)
    oracle_step(f"ragged {shape} D={D} K={K}", q, o, x)


@pytest.mark.parametrize("engine", ENGINES)
def test_empty_input(engine):
    q = vq.Quantize(64, 512, engine=engine).to(DEV).eval()
    quant, diff, ind = q(torch.empty(0, 64, device=DEV))
    assert quant.shape == (0, 64) and ind.shape == (0,) and ind.dtype == torch.int64
    assert torch.isnan(diff)            # mean over zero elements, as torch


@pytest.mark.parametrize("engine", ENGINES)
def test_embed_code_and_assign(engine):
    torch.manual_seed(1)
    q = vq.Quantize(64, 512, engine=engine).to(DEV).train()
    x = torch.randn(2, 16, 16, 64, device=DEV)
    before = {k: v.clone() for k, v in q.state_dict().items()}
    ind_only = q.assign(x)
    for k, v in q.state_dict().items():
        assert torch.equal(v, before[k])             # assign never updates the EMA buffers
    assert q.training
    q.eval()
    quant, _, ind = q(x)
    assert torch.equal(ind, ind_only)
    codes = q.embed_code(ind)                        # vqvae.py:77-78
    assert codes.shape == (2, 16, 16, 64)
    assert torch.equal(codes, q.embed.t()[ind])
    assert rel_err(quant.cpu().numpy(), (x + (codes - x)).cpu().numpy()) <= 1e-6
    q.check_ids = True
    with pytest.raises(IndexError):
        q.embed_code(torch.tensor([0, 512], device=DEV))


def test_exotic_strides_are_copied_not_rejected():
    torch.manual_seed(2)
    q = vq.Quantize(64, 128).to(DEV).eval()
    base = torch.randn(4, 8, 8, 128, device=DEV)
    x = base[..., ::2]
    quant, diff, ind = q(x)
    q2, d2, i2 = q(x.contiguous())
    assert torch.equal(ind, i2) and torch.equal(quant, q2) and torch.equal(diff, d2)


def test_gradient_formula():
    """loss = sum(quantize * w) + c * diff  =>  x.grad = w + c * 2 (x - q) / (N D)   (vqvae.py:72-73)."""
    torch.manual_seed(4)
    q = vq.Quantize(64, 512).to(DEV).train()
    for permuted in (False, True):
        x = torch.randn(2, 64, 8, 8, device=DEV).permute(0, 2, 3, 1) if permuted else torch.randn(2, 8, 8, 64, device=DEV)
        x = x.detach().requires_grad_(True)
        w = torch.randn(2, 8, 8, 64, device=DEV)
        embed_before = q.embed.clone()
        quant, diff, ind = q(x)
        (quant * w).sum().add(0.25 * diff).backward()
        codes = embed_before.t()[ind]
        expect = w + 0.25 * 2.0 * (x.detach() - codes) / x.numel()
        assert rel_err(x.grad.cpu().numpy(), expect.cpu().numpy()) <= REL_TOL
        # only diff used / only quantize used
        x2 = x.detach().clone().requires_grad_(True)
        _, d2, _ = q(x2)
        d2.backward()
        assert x2.grad is not None and torch.isfinite(x2.grad).all()


def test_multiple_forwards_per_step_see_previous_update():
    """The trainers call the same module several times per optimiser step (train_vqvae_deep.py:103-117);
    each call's outputs use the codebook as updated by the previous call."""
    torch.manual_seed(5)
    q = vq.Quantize(64, 512).to(DEV).train()
    o = QuantizeOracle(64, 512, embed=q.embed.cpu().numpy())
    for s in range(3):
        x = torch.randn(512, 64, generator=torch.Generator().manual_seed(50 + s))
        oracle_step(f"multi-forward call {s}", q, o, x.numpy())


# ---------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs 2 and 4): no CPU oracle over all rows, only
# size-independent invariants + an oracle check on a row sample
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("permuted", [False, True])
def test_full_size_properties_cfg2(engine, permuted):
    B, H, W, D, K = 128, 64, 64, 64, 512
    torch.manual_seed(0)
    q = vq.Quantize(D, K, engine=engine).to(DEV).train()
    embed0 = q.embed.clone()
    g = torch.Generator(device=DEV).manual_seed(1234)
    pick = torch.randint(0, K, (B * H * W,), device=DEV, generator=g)
    x = embed0.t()[pick] + 0.1 * torch.randn(B * H * W, D, device=DEV, generator=g)
    x = x.reshape(B, H, W, D)
    if permuted:
        x = x.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)      # NCHW-physical, as VQVAE.encode passes
    quant, diff, ind = q(x)
    assert tuple(quant.stride()) == tuple(x.stride())
    N = B * H * W
    # (1) clustered inputs return to their generating code
    assert torch.equal(ind.reshape(-1), pick)
    # (2) quantize is exactly the gather; diff is its mean squared distance
    codes = embed0.t()[ind]
    assert torch.allclose(quant, x + (codes - x), rtol=0, atol=0)
    ref_diff = ((codes - x).double() ** 2).mean()
    assert abs(float(diff) - float(ref_diff)) <= REL_TOL * float(ref_diff)
    # (3) statistics are conserved: sum(cluster_size) = (1-decay) N ; embed_avg - decay*embed0 = (1-decay) * sums
    assert abs(float(q.cluster_size.double().sum()) - 0.01 * N) <= 1e-5 * 0.01 * N
    sums = torch.zeros(K, D, device=DEV, dtype=torch.float64).index_add_(0, ind.reshape(-1), x.reshape(-1, D).double())
    expect_avg = (0.99 * embed0.double() + (1 - 0.99) * sums.t())
    assert rel_err(q.embed_avg.cpu().numpy(), expect_avg.cpu().numpy()) <= REL_TOL
    # (4) idempotence: re-quantising the codes themselves reproduces the indices with zero loss (eval, old codebook)
    q2 = vq.Quantize(D, K, engine=engine).to(DEV).eval()
    q2.embed.data.copy_(embed0)
    quant2, diff2, ind2 = q2(codes)
    assert torch.equal(ind2, ind) and float(diff2) <= 1e-10
    # (5) oracle on a row sample (reference-init codebook, fp32)
    rows = torch.randperm(N, generator=torch.Generator().manual_seed(9))[:4096]
    xs = x.reshape(-1, D)[rows.to(DEV)].cpu().numpy()
    o = QuantizeOracle(D, K, embed=embed0.cpu().numpy())
    o.training = False
    _, _, io = o.forward(xs)
    _, nbad, _ = tie_tolerant_index_mismatches(xs, embed0.cpu().numpy(), ind.reshape(-1)[rows.to(DEV)].cpu().numpy(), io)
    assert nbad == 0


@pytest.mark.parametrize("engine", ENGINES)
def test_full_size_randn_sample_vs_oracle(engine):
    """Reference-init regime at cfg-2 size: x ~ N(0,1) (many near-ties); oracle on a 16k-row sample."""
    D, K, N = 64, 512, 128 * 64 * 64
    torch.manual_seed(0)
    q = vq.Quantize(D, K, engine=engine).to(DEV).eval()
    x = torch.randn(N, D, device=DEV, generator=torch.Generator(device=DEV).manual_seed(77))
    ind = q.assign(x)
    rows = torch.arange(0, N, N // 16384, device=DEV)[:16384]
    xs = x[rows].cpu().numpy()
    e = q.embed.cpu().numpy()
    o = QuantizeOracle(D, K, embed=e)
    o.training = False
    _, _, io = o.forward(xs)
    ndiff, nbad, _ = tie_tolerant_index_mismatches(xs, e, ind[rows].cpu().numpy(), io)
    assert nbad == 0, f"{nbad} of 16384 sampled rows disagree beyond fp32 near-ties"


def test_host_buffer_path_matches_device_path():
    lib = _native.load()
    D, K, N = 64, 512, 70000
    torch.manual_seed(0)
    q = vq.Quantize(D, K).to(DEV).train()
    q2 = vq.Quantize(D, K).to(DEV).train()
    q2.load_state_dict(q.state_dict())
    x_h = torch.randn(N, D).pin_memory()
    quant_d, diff_d, ind_d = q(x_h.to(DEV))
    ctx = C.c_void_p()
    _native.check(lib.vqb200_host_ctx_create(N, D, K, C.byref(ctx)), "ctx")
    try:
        q_h = torch.empty(N, D).pin_memory()
        i_h = torch.empty(N, dtype=torch.int64).pin_memory()
        d_h = torch.empty(1).pin_memory()
        torch.cuda.synchronize()
        _native.check(lib.vqb200_host_quantize(ctx, C.c_void_p(x_h.data_ptr()), N, _native.ptr(q2.embed),
                                               _native.ptr(q2.cluster_size), _native.ptr(q2.embed_avg), 0.99,
                                               float(1 - 0.99), 1e-5, 1, C.c_void_p(q_h.data_ptr()),
                                               C.c_void_p(i_h.data_ptr()), C.c_void_p(d_h.data_ptr()), 0), "host_quantize")
    finally:
        lib.vqb200_host_ctx_destroy(ctx)
    assert torch.equal(i_h, ind_d.cpu())
    assert torch.equal(q_h, quant_d.cpu())
    assert abs(float(d_h) - float(diff_d)) <= 1e-6 * float(diff_d)
    assert rel_err(q2.cluster_size.cpu().numpy(), q.cluster_size.cpu().numpy()) <= 1e-6
    assert col_rel_err(q2.embed_avg.cpu().numpy(), q.embed_avg.cpu().numpy()) <= REL_TOL
    assert col_rel_err(q2.embed.cpu().numpy(), q.embed.cpu().numpy()) <= REL_TOL


def test_host_buffer_path_data_parallel_form():
    """vqb200_host_quantize_stats (statistics left on the device, no EMA) + vqb200_ema_update == vqb200_host_quantize with
    training = 1: the split a data-parallel caller uses to all-reduce the statistics in between (vqvae.py:58-59)."""
    lib = _native.load()
    D, K, N = 64, 512, 50000
    torch.manual_seed(4)
    q = vq.Quantize(D, K).to(DEV).train()
    q2 = vq.Quantize(D, K).to(DEV).train()
    q2.load_state_dict(q.state_dict())
    x_h = torch.randn(N, D).pin_memory()
    stats = torch.zeros(lib.vqb200_stats_bytes(D, K) // 4, device=DEV)
    ctx = C.c_void_p()
    _native.check(lib.vqb200_host_ctx_create(N, D, K, C.byref(ctx)), "ctx")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    try:
        out = [(torch.empty(N, D).pin_memory(), torch.empty(N, dtype=torch.int64).pin_memory(), torch.empty(1).pin_memory())
               for _ in range(2)]
        torch.cuda.synchronize()
        for step in range(2):                                    # two steps: the second sees the first one's EMA update
            qh, ih, dh = out[0]
            _native.check(lib.vqb200_host_quantize(ctx, C.c_void_p(x_h.data_ptr()), N, _native.ptr(q.embed),
                                                   _native.ptr(q.cluster_size), _native.ptr(q.embed_avg), 0.99,
                                                   float(1 - 0.99), 1e-5, 1, C.c_void_p(qh.data_ptr()),
                                                   C.c_void_p(ih.data_ptr()), C.c_void_p(dh.data_ptr()), 0), "host_quantize")
            qs, is_, ds = out[1]
            _native.check(lib.vqb200_host_quantize_stats(ctx, C.c_void_p(x_h.data_ptr()), N, _native.ptr(q2.embed),
                                                         _native.ptr(stats), C.c_void_p(qs.data_ptr()),
                                                         C.c_void_p(is_.data_ptr()), C.c_void_p(ds.data_ptr()), 0), "host_quantize_stats")
            assert abs(float(stats[K * D: K * D + K].sum()) - N) < 0.5
            _native.check(lib.vqb200_ema_update(_native.ptr(stats), _native.ptr(q2.cluster_size), _native.ptr(q2.embed_avg),
                                                _native.ptr(q2.embed), D, K, 0.99, float(1 - 0.99), 1e-5, None, st), "ema_update")
            torch.cuda.synchronize()
            if step == 0:
                assert torch.equal(ih, is_) and torch.equal(qh, qs)
            else:
                assert float((ih != is_).float().mean()) <= 1e-4
            assert abs(float(dh) - float(ds)) <= 1e-5 * float(dh)
            assert rel_err(q2.cluster_size.cpu().numpy(), q.cluster_size.cpu().numpy()) <= 1e-6
            assert col_rel_err(q2.embed_avg.cpu().numpy(), q.embed_avg.cpu().numpy()) <= REL_TOL
            assert col_rel_err(q2.embed.cpu().numpy(), q.embed.cpu().numpy()) <= REL_TOL
    finally:
        lib.vqb200_host_ctx_destroy(ctx)
