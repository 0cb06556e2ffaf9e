"""SURVEY 8f row 4 on two GPUs: the reference trainer's step (train_vqvae.py:85-118,166-171: DDP with per-forward buffer
broadcast, `recon_loss.item()` and the reference's pickled `all_gather` every step) against the same step with
`ddp_wrap` (no buffer broadcast) + `DeferredMetrics` (no per-step synchronisation) -- the unmodified reference VQVAE with
only the `Quantize` class swapped, B = 8 per GPU at 256 px (the reference's real batch size).  Checks that both loops report
the same running MSE, that the replicas stay bit-identical without the broadcast, and writes the step times to
gpurun_out/trainer_overlap.json.  Needs 2 GPUs (skipped on a 1-GPU box)."""
import json
import os
import socket
import time

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import vq_vae_2_pytorch_b200 as vq
        from oracle import reference_module
        ref = reference_module.load("vqvae")
        import distributed as ref_dist                           # the reference's package (oracle/_ref on sys.path)
        orig = ref.Quantize
        ref.Quantize = vq.Quantize
        try:
            torch.manual_seed(0)
            model_a = ref.VQVAE().to(dev).train()
            torch.manual_seed(0)
            model_b = ref.VQVAE().to(dev).train()
        finally:
            ref.Quantize = orig
        imgs = [torch.randn(8, 3, 256, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(100 * rank + i)) for i in range(4)]
        res = {}
        # ---- A: as the reference trainer does it
        # (the fork's VQVAE carries an unused `dec_ir` decoder, vqvae.py:203-210: DDP needs find_unused_parameters for it)
        ddp = torch.nn.parallel.DistributedDataParallel(model_a, device_ids=[rank], output_device=rank,
                                                        find_unused_parameters=True)                      # train_vqvae.py:166-171
        opt = torch.optim.Adam(ddp.parameters(), lr=3e-4)

        def step_ref(img, state):
            ddp.zero_grad()
            o, latent = ddp(img)
            recon = (o - img).pow(2).mean()
            (recon + 0.25 * latent.mean()).backward()
            opt.step()
            comm = {"mse_sum": recon.item() * img.shape[0], "mse_n": img.shape[0]}        # train_vqvae.py:93-100
            for part in ref_dist.all_gather(comm):
                state[0] += part["mse_sum"]; state[1] += part["mse_n"]

        # ---- B: the recipe
        ddp2 = vq.ddp_wrap(model_b, dev, find_unused_parameters=True)
        opt2 = torch.optim.Adam(ddp2.parameters(), lr=3e-4)
        metrics = vq.DeferredMetrics(dev, ("mse_sum", "mse_n"))

        def step_new(img, state):
            ddp2.zero_grad()
            o, latent = ddp2(img)
            recon = (o - img).pow(2).mean()
            (recon + 0.25 * latent.mean()).backward()
            opt2.step()
            metrics.add(mse_sum=recon.detach() * img.shape[0], mse_n=img.shape[0])

        for name, fn in (("reference_style", step_ref), ("recipe", step_new)):
            state = [0.0, 0]
            for i in range(4):
                fn(imgs[i % 4], [0.0, 0])
            metrics.reset()
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            for i in range(24):
                fn(imgs[i % 4], state)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[name] = {"ms_per_step": float(t.item()) / 24 * 1e3}
            if name == "reference_style":
                res[name]["mse"] = state[0] / state[1]
            else:
                tot = metrics.totals()
                res[name]["mse"] = tot["mse_sum"] / tot["mse_n"]
        res["replicas_identical_without_buffer_broadcast"] = vq.replicas_identical(model_b)
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_two_gpu_trainer_step_reference_style_vs_recipe():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import reference_module
    if not reference_module.available():
        pytest.skip("reference not staged")
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = dict(out)[0]
    assert res["replicas_identical_without_buffer_broadcast"]
    # both loops ran the same model on the same images from the same seed: the same running MSE (their parameters follow
    # the same trajectory up to all-reduce summation order)
    assert abs(res["recipe"]["mse"] - res["reference_style"]["mse"]) <= 1e-3 * abs(res["reference_style"]["mse"])
    assert res["recipe"]["ms_per_step"] <= 1.05 * res["reference_style"]["ms_per_step"]
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "trainer_overlap.json"), "w") as f:
        json.dump({"what": "unmodified reference VQVAE with the Quantize class swapped, 2 x B200, B = 8 per GPU, 256 px, Adam, 24 steps",
                   **res}, f, indent=1)
