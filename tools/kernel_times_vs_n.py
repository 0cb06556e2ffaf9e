"""Per-kernel device durations (CUPTI via torch.profiler) of one training forward at several batch sizes: does the statistics
kernel speed up when x is still L2-resident (N * 256 B well below the 126 MB L2)?"""
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402

dev = "cuda:0"
D, K = 64, 512
torch.manual_seed(0)
for N in (32768, 65536, 131072, 262144, 524288):
    q = vq.Quantize(D, K).to(dev).train()
    xs = []
    for i in range(3):
        pick = torch.randint(0, K, (N,), device=dev)
        xs.append((q.embed.t()[pick] + 0.1 * torch.randn(N, D, device=dev)).contiguous())
    q.cluster_size.data.fill_(N / K); q.embed_avg.data.copy_(q.embed * (N / K))
    for i in range(5):
        q(xs[i % 3])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(12):
            q(xs[i % 3])
        torch.cuda.synchronize()
    acc = defaultdict(list)
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            acc[e.name.split("(")[0][-28:]].append(e.time_range.end - e.time_range.start)
    line = "  ".join(f"{k}: {sorted(v)[len(v) // 2]:.1f}" for k, v in sorted(acc.items()))
    print(f"N={N:7d} ({N * 256 / 1e6:5.1f} MB of x): {line}", flush=True)
