"""Collective row (vqvae.py:58-59 -> distributed.py:64-72) on ONE GPU: the all-reduce that is fused
into the EMA kernel over peer memory (`vqb200_ema_update_p2p`) driven with world = 1 -- its own statistics buffer and
flag array stand for the peer-mapped ones -- must reproduce `vqb200_ema_update` bit for bit over several steps (flag
publication, wait, rank-ordered sum of one rank, step/parity bookkeeping).  The multi-rank run of the same kernel is
checked by bench.py at N > 1 ("parity_multi") and by tools/p2p_check.py; the packing / reduction algebra by the gloo test.
"""
import ctypes as C

import pytest
import torch

import vq_vae_2_pytorch_b200 as vq
from vq_vae_2_pytorch_b200 import _native

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("K", [512, 256])
def test_p2p_ema_with_world_1_equals_local_ema(K):
    lib = _native.load()
    D = 64
    torch.manual_seed(0)
    a = vq.Quantize(D, K).to(DEV).train()
    b = vq.Quantize(D, K).to(DEV).train()
    b.load_state_dict(a.state_dict())
    n = lib.vqb200_stats_bytes(D, K) // 4
    n_al = (n + 63) // 64 * 64
    buf = torch.zeros(2 * n_al + 128, device=DEV)                # [stats parity 0 | stats parity 1 | flags 0 | flags 1]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for step in range(1, 5):
        x = torch.randn(128 * 37 + 5, D, device=DEV, generator=torch.Generator(device=DEV).manual_seed(step))
        par = step & 1
        stats = buf[par * n_al: par * n_al + n]
        ws = b._workspace(x.device, x.shape[0])
        quant = torch.empty_like(x); ind = torch.empty(x.shape[0], dtype=torch.int64, device=DEV); diff = torch.empty((), device=DEV)
        # ONE forward (no EMA) writes this step's packed statistics; both EMA variants then consume the same bits (the
        # statistics kernel's summation order is not reproducible between two runs, like the reference's GEMM)
        _native.check(lib.vqb200_quantize_step(x.data_ptr(), x.shape[0], D, K, x.shape[0], 0, D, 1, b.embed.data_ptr(),
                                               b.cluster_size.data_ptr(), b.embed_avg.data_ptr(), ws["image"].data_ptr(),
                                               quant.data_ptr(), ind.data_ptr(), diff.data_ptr(), stats.data_ptr(),
                                               ws["scratch"].data_ptr(), None, 0, 0, 0.99, float(1 - 0.99), 1e-5, st), "step")
        local_stats = stats.clone()
        _native.check(lib.vqb200_ema_update(local_stats.data_ptr(), a.cluster_size.data_ptr(), a.embed_avg.data_ptr(),
                                            a.embed.data_ptr(), D, K, 0.99, float(1 - 0.99), 1e-5, None, st), "ema_local")
        stats_ptrs = (C.c_void_p * 1)(stats.data_ptr())
        flag_ptrs = (C.c_void_p * 1)(buf.data_ptr() + 4 * (2 * n_al + 64 * par))
        _native.check(lib.vqb200_ema_update_p2p(stats_ptrs, flag_ptrs, 0, 1, step, b.cluster_size.data_ptr(),
                                                b.embed_avg.data_ptr(), b.embed.data_ptr(), D, K, 0.99, float(1 - 0.99), 1e-5,
                                                None, st), "ema_p2p")
        torch.cuda.synchronize()
        for name in ("cluster_size", "embed_avg", "embed"):
            assert torch.equal(getattr(a, name), getattr(b, name)), f"step {step}: {name} differs between the fused P2P and the local EMA"
        assert int(buf[2 * n_al + 64 * par: 2 * n_al + 64 * par + 1].view(torch.int32)) == step      # the flag this rank published
    # four EMA steps of 4741 rows each from cluster_size = 0:  sum = N (1-d) (d^3 + d^2 + d + 1)
    want = (128 * 37 + 5) * 0.01 * (0.99 ** 3 + 0.99 ** 2 + 0.99 + 1)
    assert abs(float(a.cluster_size.double().sum()) - want) <= 1e-5 * want


def test_p2p_ema_rejects_bad_arguments():
    lib = _native.load()
    z = torch.zeros(64 * 513 + 4 + 128, device=DEV)
    ptrs = (C.c_void_p * 1)(z.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.vqb200_ema_update_p2p(ptrs, ptrs, 0, 1, 0, z.data_ptr(), z.data_ptr(), z.data_ptr(), 64, 512, 0.99, 0.01, 1e-5, None, st) == -1   # step 0
    assert lib.vqb200_ema_update_p2p(ptrs, ptrs, 1, 1, 1, z.data_ptr(), z.data_ptr(), z.data_ptr(), 64, 512, 0.99, 0.01, 1e-5, None, st) == -1   # rank >= world
    assert lib.vqb200_ema_update_p2p(ptrs, ptrs, 0, 1, 1, z.data_ptr(), z.data_ptr(), z.data_ptr(), 48, 100, 0.99, 0.01, 1e-5, None, st) == -2   # shape


@pytest.mark.parametrize("K", [512, 256])
def test_step_peers_with_world_1_equals_single_rank_step(K):
    """What the module runs at world > 1 (vqb200_quantize_step_peers: forward, then ONE kernel that folds the statistics
    tables, stores them as {value, step} pairs into every rank's receive slot, polls its local slots and applies the EMA to
    the rank-ordered sum) driven with world = 1 against the single-rank step on the same inputs: identical indices /
    outputs, EMA buffers equal up to the statistics' summation-order noise, every word of the slot tagged with the step,
    the time-out words clear, and the slot holding the statistics (counts sum to the rows)."""
    lib = _native.load()
    D = 64
    torch.manual_seed(1)
    a = vq.Quantize(D, K).to(DEV).train()
    b = vq.Quantize(D, K).to(DEV).train()
    b.load_state_dict(a.state_dict())
    n = K * (D + 1)
    n_al = (n + 63) // 64 * 64
    buf = torch.zeros(2 * 2 * n_al + 64, device=DEV)             # [slot parity 0 | slot parity 1 | err (2) .. step counter (+32)]
    err = buf[4 * n_al: 4 * n_al + 64]
    slots = [buf[par * 2 * n_al: par * 2 * n_al + 2 * n] for par in (0, 1)]
    dst = (C.c_void_p * 2)(*[t.data_ptr() for t in slots]); recv = (C.c_void_p * 2)(*[t.data_ptr() for t in slots])
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    from vq_vae_2_pytorch_b200 import row_layout
    for step in range(1, 6):
        g = torch.Generator(device=DEV).manual_seed(step)
        x = torch.randn(2, D, 32, 32, device=DEV, generator=g).permute(0, 2, 3, 1) if step % 2 else torch.randn(128 * 29 + 3, D, device=DEV, generator=g)
        qa, da, ia = a(x)                                        # single-rank step (fold fused into the EMA kernel as well)
        nr, rpi, img, row, col = row_layout(x)
        slot = slots[step & 1]                                   # the kernel's own counter picks the parity: step = counter + 1
        ws = b._workspace(x.device, nr)
        quant = torch.empty_strided(x.shape, x.stride(), device=DEV)
        ind = torch.empty(x.shape[:-1], dtype=torch.int64, device=DEV); diff = torch.empty((), device=DEV)
        _native.check(lib.vqb200_quantize_step_peers(x.data_ptr(), nr, D, K, rpi, img, row, col, b.embed.data_ptr(),
                                                     b.cluster_size.data_ptr(), b.embed_avg.data_ptr(), ws["image"].data_ptr(),
                                                     quant.data_ptr(), ind.data_ptr(), diff.data_ptr(), ws["scratch"].data_ptr(), None, 0,
                                                     0.99, float(1 - 0.99), 1e-5, dst, recv, err.data_ptr(), err[32:].data_ptr(), 0, 1, st), "step_peers")
        torch.cuda.synchronize()
        assert torch.equal(ind, ia) and torch.equal(quant, qa) and abs(float(diff) - float(da)) <= 1e-6 * abs(float(da))
        pairs = slot.view(n, 2)
        assert bool((pairs[:, 1].view(torch.int32) == step).all())          # every word carries the step tag
        assert int(err[:1].view(torch.int32)) == 0                          # no time-out recorded
        assert int(err[32:33].view(torch.int32)) == step                    # the device-side exchange counter advanced
        vals = pairs[:, 0]
        assert abs(float(vals[K * D: K * D + K].sum()) - nr) < 0.5          # the pushed counts
        assert torch.allclose(vals[: K * D].view(K, D).sum(0), x.reshape(-1, D).sum(0), rtol=1e-4, atol=1e-2)
        assert torch.equal(a.cluster_size, b.cluster_size)                  # integer counts: exact
        scale = a.embed_avg.abs().amax(0, keepdim=True).clamp_min(1e-30)
        assert float(((a.embed_avg - b.embed_avg).abs() / scale).max()) <= 2e-6
        b.load_state_dict(a.state_dict())


def test_stats_exchange_with_world_1_is_the_identity_and_counts_steps():
    """vqb200_stats_exchange_peers (the in-place all-reduce over peer memory that shapes outside the fused kernel use) with
    world = 1: the rank-ordered sum over one rank is the buffer itself; every slot word carries the step tag; the device-side
    counter advances by one per launch and the ticket word returns to zero; parity alternates between the two slots."""
    lib = _native.load()
    n = 256 * 257 + 3                                            # not a multiple of anything
    n_al = (n + 63) // 64 * 64
    buf = torch.zeros(2 * 2 * n_al + 64, device=DEV)
    slots = [buf[par * 2 * n_al: par * 2 * n_al + 2 * n] for par in (0, 1)]
    err = buf[4 * n_al: 4 * n_al + 64]
    ptrs = (C.c_void_p * 2)(*[t.data_ptr() for t in slots])
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for step in range(1, 5):
        stats = torch.randn(n, device=DEV, generator=torch.Generator(device=DEV).manual_seed(step))
        want = stats.clone()
        _native.check(lib.vqb200_stats_exchange_peers(stats.data_ptr(), n, ptrs, ptrs, err.data_ptr(), err[32:].data_ptr(), 0, 1, st),
                      "stats_exchange")
        torch.cuda.synchronize()
        assert torch.equal(stats, want)
        pairs = slots[step & 1].view(n, 2)
        assert torch.equal(pairs[:, 0], want) and bool((pairs[:, 1].view(torch.int32) == step).all())
        words = err.view(torch.int32)
        assert int(words[0]) == 0 and int(words[32]) == step and int(words[33]) == 0
    assert lib.vqb200_stats_exchange_peers(None, n, ptrs, ptrs, err.data_ptr(), err[32:].data_ptr(), 0, 1, st) == -1
    assert lib.vqb200_stats_exchange_peers(buf.data_ptr(), n, ptrs, ptrs, err.data_ptr(), err[32:].data_ptr(), 1, 1, st) == -1
