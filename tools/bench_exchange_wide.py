"""Data-parallel training step of shapes OUTSIDE the fused fold + EMA kernel: in-place peer-memory exchange kernel
(vqb200_stats_exchange_peers) against one NCCL all-reduce of the packed statistics.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_exchange_wide.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
for D, K, N in ((256, 512, 131072), (256, 512, 524288), (128, 512, 524288), (64, 1024, 524288), (64, 2048, 131072)):
    res = {}
    for name, no_p2p in (("peer memory", False), ("nccl", True)):
        torch.manual_seed(0)
        q = vq.Quantize(D, K).to(dev).train()
        xs = []
        for i in range(3):
            g = torch.Generator(device=dev).manual_seed(100 * rank + i)
            pick = torch.randint(0, K, (N,), device=dev, generator=g)
            xs.append((q.embed.t()[pick] + 0.1 * torch.randn(N, D, device=dev, generator=g)).contiguous())
        q.cluster_size.data.fill_(world * N / K); q.embed_avg.data.copy_(q.embed * (world * N / K))
        if no_p2p:
            os.environ["VQB200_NO_P2P"] = "1"
        for i in range(5):
            q(xs[i % 3])
        os.environ.pop("VQB200_NO_P2P", None)
        took = "peer memory" if q._ws[dev]["peer"] is not None else "nccl"
        assert took == name, (took, name)
        times = []
        for w in range(5):
            dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(20):
                q(xs[i % 3])
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / 20 * 1e3], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t))
        res[name] = sorted(times)[2]
    if rank == 0:
        print(f"D={D:3d} K={K:4d} N={N:6d} x {world} ranks: peer-memory exchange {res['peer memory']:7.1f} us/step, NCCL all-reduce {res['nccl']:7.1f} us/step", flush=True)
dist.destroy_process_group()
