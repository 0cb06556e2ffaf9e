"""Device timeline (CUPTI via torch.profiler) of steady-state training steps at cfg-2: start / end of every kernel relative to the
first, so that gaps between the five launches of a step are visible."""
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402

dev = "cuda:0"
D, K, N = 64, 512, 524288
torch.manual_seed(0)
q = vq.Quantize(D, K).to(dev).train()
xs = []
for i in range(3):
    pick = torch.randint(0, K, (N,), device=dev)
    xs.append((q.embed.t()[pick] + 0.1 * torch.randn(N, D, device=dev)).contiguous())
q.cluster_size.data.fill_(N / K); q.embed_avg.data.copy_(q.embed * (N / K))
for i in range(10):
    q(xs[i % 3])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(10):
        q(xs[i % 3])
    torch.cuda.synchronize()
evs = sorted((e.time_range.start, e.time_range.end, e.name.split("(")[0][-26:]) for e in prof.events()
             if e.device_type == torch.autograd.DeviceType.CUDA)
t0 = evs[0][0]
prev_end = None
for a, b, n in evs[20:36]:
    gap = "" if prev_end is None else f"  gap to previous end {a - prev_end:+6.1f}"
    print(f"{a - t0:9.1f} {b - t0:9.1f}  {b - a:6.1f} us  {n}{gap}")
    prev_end = b
mains = [a for a, b, n in evs if "k_vq_tc" in n]
print("step period (main kernel start to start):", [round(y - x, 1) for x, y in zip(mains, mains[1:])])
