"""Where a multi-rank step spends its time inside the fold + exchange + EMA kernel (needs a library built with
-DVQB200_P2P_TRACE; run under torchrun).  Prints per rank the mean of (fold, push + flag stores, wait for the peers, EMA) in us
of the kernel's last block, plus the step time, with back-to-back steps (trace words read after the loop: last step only per
parity) and with a device sync after every step."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
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
peer = q._ws[dev]["peer"]
assert peer is not None
slots = (peer["buf"].numel() - 64) // 2


def trace():
    return peer["buf"][2 * slots + 4: 2 * slots + 10].view(torch.int32).cpu().tolist()


rows = []
for i in range(40):
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    q(xs[i % 3])
    b.record()
    torch.cuda.synchronize()
    t = trace()
    rows.append([a.elapsed_time(b) * 1e3] + [v / 1e3 for v in t[1:5]])
m = torch.tensor(rows[5:]).mean(0).tolist()
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f"rank {rank}: isolated steps: step {m[0]:.1f} us | fold {m[1]:.2f}  push {m[2]:.2f}  poll+sum {m[3]:.2f}  ema {m[4]:.2f} us", flush=True)
dist.destroy_process_group()
