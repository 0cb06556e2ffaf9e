"""SURVEY 8f row 4 (trainer-level overlap): `DeferredMetrics` + `ddp_wrap` against what the reference trainers do
(train_vqvae.py:93-118: `recon_loss.item()` and a pickled `all_gather` every step; :166-171: DDP with buffer broadcast).

CPU (gloo, world_size 2): the deferred device-side sums equal the reference's per-step pickled all_gather aggregate (the
reference's own `distributed.all_gather`, staged unmodified in oracle/_ref, is restated with all_gather_object on CPU: the original needs CUDA), and DDP built
by `ddp_wrap` does not broadcast buffers.
GPU: a training loop of the class-swapped reference VQVAE with `DeferredMetrics` runs with implicit host synchronisations
turned into errors (torch.cuda.set_sync_debug_mode), i.e. the quantizer and the glue never force a sync in the step.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vq_vae_2_pytorch_b200 as vq
from oracle import reference_module


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the reference's pickled all_gather (distributed.py:75-107) moves its byte tensors to "cuda", so on CPU its semantics
        # are executed through torch's all_gather_object (same pickle -> gather -> unpickle on every rank)
        def ref_all_gather(data):
            parts = [None] * world
            dist.all_gather_object(parts, data)
            return parts
        m = vq.DeferredMetrics("cpu", ("mse_sum", "mse_n"))
        mse_sum = mse_n = 0.0
        g = torch.Generator().manual_seed(7 + rank)
        for step in range(6):
            n = 4 + rank                                         # ragged last batches differ per rank
            recon = torch.rand((), generator=g)
            m.add(mse_sum=recon * n, mse_n=n)
            comm = {"mse_sum": recon.item() * n, "mse_n": n}     # train_vqvae.py:93-100
            comm = ref_all_gather(comm)
            for part in comm:
                mse_sum += part["mse_sum"]
                mse_n += part["mse_n"]
        tot = m.totals()
        assert abs(tot["mse_sum"] - mse_sum) <= 1e-6 * mse_sum and tot["mse_n"] == mse_n
        assert tot["mse_n"] == 6 * (4 + 5)
        # DDP without the per-forward buffer broadcast
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.BatchNorm1d(4))
        ddp = vq.ddp_wrap(net)
        assert ddp.broadcast_buffers is False
        assert vq.replicas_identical(net)                        # no Quantize inside: trivially true
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_deferred_metrics_equal_the_reference_pickled_all_gather():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_deferred_metrics_single_process():
    m = vq.DeferredMetrics("cpu", ("a", "b"))
    m.add(a=torch.tensor(1.5), b=2)
    m.add(a=0.5)
    assert m.totals() == {"a": 2.0, "b": 2.0}
    m.reset()
    assert m.totals() == {"a": 0.0, "b": 0.0}


@pytest.mark.gpu
def test_training_loop_has_no_implicit_synchronisation():
    try:
        ref = reference_module.load("vqvae")
    except reference_module.ReferenceUnavailable as exc:
        pytest.skip(str(exc))
    dev = torch.device("cuda:0")
    orig = ref.Quantize
    ref.Quantize = vq.Quantize
    try:
        torch.manual_seed(0)
        model = ref.VQVAE().to(dev).train()
    finally:
        ref.Quantize = orig
    opt = torch.optim.Adam(model.parameters(), lr=3e-4)
    metrics = vq.DeferredMetrics(dev, ("mse_sum", "mse_n"))
    imgs = [torch.randn(4, 3, 256, 256, device=dev) for _ in range(3)]
    for img in imgs[:2]:                                         # warm-up: allocator, cuDNN autotuning, workspaces
        out, latent = model(img)
        ((out - img).pow(2).mean() + 0.25 * latent.mean()).backward()
        opt.step(); opt.zero_grad()
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        for img in imgs * 2:
            out, latent = model(img)                             # train_vqvae.py:85-91
            recon = (out - img).pow(2).mean()
            (recon + 0.25 * latent.mean()).backward()
            opt.step(); opt.zero_grad()
            metrics.add(mse_sum=recon.detach() * img.shape[0], mse_n=img.shape[0])
    finally:
        torch.cuda.set_sync_debug_mode("default")
    tot = metrics.totals()
    assert tot["mse_n"] == 24 and tot["mse_sum"] > 0
