"""Per-kernel device times of real training steps (CUPTI activity trace via torch.profiler: back-to-back launches,
warm caches -- unlike the ncu launch list, which serialises and flushes).  Usage: python tools/kernel_trace.py [engine] [dist]"""
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else "auto"
dist = sys.argv[2] if len(sys.argv) > 2 else "clustered"
layout = sys.argv[3] if len(sys.argv) > 3 else "dense"
B, H, W, D, K = 128, 64, 64, int(os.environ.get("VQ_TRACE_D", "64")), int(os.environ.get("VQ_TRACE_K", "512"))
N = B * H * W
import os
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
dev = f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as tdist
    tdist.init_process_group("nccl", device_id=torch.device(dev))
torch.manual_seed(0)
q = vq.Quantize(D, K, engine=engine).to(dev).train()
if os.environ.get("VQ_TRACE_EVAL"):
    q.eval()
embed0 = q.embed.clone()
xs = []
for i in range(3):
    g = torch.Generator(device=dev).manual_seed(1234 + 1000 * i + rank)
    if dist == "clustered":
        pick = torch.randint(0, K, (N,), device=dev, generator=g)
        x = embed0.t()[pick] + 0.1 * torch.randn(N, D, device=dev, generator=g)
    else:
        x = torch.randn(N, D, device=dev, generator=g)
    x = x.reshape(B, H, W, D).contiguous()
    if layout == "nchw":
        x = x.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    xs.append(x)
if dist == "clustered":
    q.cluster_size.data.fill_(float(world * N) / K)
    q.embed_avg.data.copy_(embed0 * (float(world * N) / K))
for i in range(10):
    q(xs[i % 3])
torch.cuda.synchronize()
steps = 20
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        q(xs[i % 3])
    torch.cuda.synchronize()
tot = defaultdict(float)
cnt = defaultdict(int)
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        tot[ev.name[:70]] += ev.device_time
        cnt[ev.name[:70]] += 1
total = sum(tot.values())
if rank != 0:
    sys.exit(0)
print(f"engine={engine} dist={dist}: {total / steps:.1f} us of kernel time per step")
for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{t / steps:9.2f} us/step  x{cnt[name] / steps:4.1f}  {name}")
