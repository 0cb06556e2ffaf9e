"""Per-step device time of the training forward on every rank (CUDA events between steps), to see where multi-rank
steps lose time.  Run under torchrun; VQB200_NO_P2P=1 selects the NCCL path."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
D, K, N = 64, 512, 128 * 64 * 64
q = vq.Quantize(D, K).to(dev).train()
e0 = q.embed.clone()
xs = []
for i in range(3):
    g = torch.Generator(device=dev).manual_seed(1234 + 1000 * i + rank)
    pick = torch.randint(0, K, (N,), device=dev, generator=g)
    xs.append((e0.t()[pick] + 0.1 * torch.randn(N, D, device=dev, generator=g)).reshape(128, 64, 64, D))
q.cluster_size.data.fill_(float(world * N) / K)
q.embed_avg.data.copy_(e0 * (float(world * N) / K))
for i in range(12):
    q(xs[i % 3])
torch.cuda.synchronize()
steps = 24
evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
evs[0].record()
for i in range(steps):
    q(xs[i % 3])
    evs[i + 1].record()
torch.cuda.synchronize()
d = [evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(steps)]
for r in range(world):
    if world > 1:
        dist.barrier()
    if r == rank:
        print(f"[rank {rank}] us/step: " + " ".join(f"{v:.0f}" for v in d) + f" | mean {sum(d) / len(d):.1f}", flush=True)
pw = q._ws.get(dev, {}).get("peer")
if pw is not None and os.environ.get("VQB200_P2P_TRACE"):
    n_al = (len(pw["stats"][0]) + 63) // 64 * 64
    fl = pw["buf"][2 * n_al:2 * n_al + 128].view(torch.int32).cpu().tolist()
    recs = []
    for par in (0, 1):
        for i in range(14):
            t0, w, st = fl[64 * par + 8 + 4 * i: 64 * par + 8 + 4 * i + 3]
            if st:
                recs.append((st & 0xffffffff, t0 & 0xffffffff, w))
    recs.sort()
    out = []
    for (s0, a, w0), (s1, b, w1) in zip(recs[:-1], recs[1:]):
        out.append(f"{s1}:{((b - a) & 0xffffffff) / 1e3:.0f}/{w1 / 1e3:.0f}")
    for r in range(world):
        dist.barrier()
        if r == rank:
            print(f"[rank {rank}] step:dt_us/wait_us " + " ".join(out[-20:]), flush=True)
if world > 1:
    dist.destroy_process_group()
