"""Two GPUs: the multi-rank training forward (vqb200_quantize_step_peers -- vqvae.py:42-70 with the all-reduce of
distributed/distributed.py:64-72 fused into the EMA kernel) captured ONCE in a CUDA graph and replayed.  The step tag of the
exchange is a device-side counter owned by the kernel, so a replay sends fresh tags; checks after every replay that no word
timed out, that the counter advanced, that the replicas are bit-identical, and that the buffers follow the same
trajectory as an eager twin on the NCCL all-reduce path.  Needs 2 GPUs (skipped on a 1-GPU box)."""
import os
import socket

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
        from vq_vae_2_pytorch_b200 import replicas_identical
        torch.manual_seed(3)
        a = vq.Quantize(64, 512).to(dev).train()
        b = vq.Quantize(64, 512).to(dev).train()
        e0 = a.embed.clone()
        xs = []
        for i in range(3):
            g = torch.Generator(device=dev).manual_seed(1000 * rank + i)
            pick = torch.randint(0, 512, (8 * 32 * 32,), device=dev, generator=g)
            xs.append((e0.t()[pick] + 0.3 * torch.randn(pick.numel(), 64, device=dev, generator=g)).reshape(8, 32, 32, 64))
        state0 = {k: v.clone() for k, v in a.state_dict().items()}
        a(xs[0])                                      # collective set-up of the peer workspace, outside the capture
        os.environ["VQB200_NO_P2P"] = "1"             # the twin takes the NCCL all-reduce + separate EMA path
        b(xs[0])
        del os.environ["VQB200_NO_P2P"]
        peer = a._ws[dev]["peer"]
        res = {"peer_path": peer is not None, "twin_on_nccl": b._ws[dev]["peer"] is None}
        if peer is None:
            res["why"] = getattr(a, "_peer_error", "?")
            out[rank] = res
            return
        a.load_state_dict(state0)
        b.load_state_dict(state0)
        torch.cuda.synchronize()
        dist.barrier()
        slots = (peer["buf"].numel() - 64) // 2
        counter = peer["buf"][2 * slots + 32: 2 * slots + 33].view(torch.int32)
        c0 = int(counter)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            outs = [a(x) for x in xs]
        a.load_state_dict(state0)
        ok_ind, ok_traj, ok_rep, ok_cnt, ok_err = True, True, True, True, True
        for replay in range(4):
            graph.replay()
            torch.cuda.synchronize()
            ok_cnt &= int(counter) == c0 + 3 * (replay + 1)
            ok_err &= int(peer["err"][0]) == 0
            for x, (quant, diff, ind) in zip(xs, outs):
                eq, ed, ei = b(x)
                # same trajectory up to the summation order of the statistics: indices may differ only at near-ties
                ok_ind &= float((ind != ei).float().mean()) <= 1e-4
            ok_rep &= replicas_identical(a)
            for name in ("cluster_size", "embed_avg", "embed"):
                ga, gb = getattr(a, name), getattr(b, name)
                ok_traj &= bool(torch.allclose(ga, gb, rtol=1e-4, atol=1e-5 * float(gb.abs().max())))
        res.update(indices=ok_ind, trajectory=ok_traj, replicas=ok_rep, counter=ok_cnt, no_timeout=ok_err,
                   exchanges=int(counter) - c0)
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_multi_rank_step_replays_from_a_cuda_graph():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for rank in (0, 1):
        res = dict(out)[rank]
        assert res["peer_path"], res
        assert res["twin_on_nccl"]
        assert res["no_timeout"] and res["counter"] and res["exchanges"] == 12, res
        assert res["replicas"] and res["indices"] and res["trajectory"], res
