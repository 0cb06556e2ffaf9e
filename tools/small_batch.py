"""Step time of the training forward at the small batches the reference's trainers really use (B = 8 .. 32 per GPU):
eager calls (host-issue bound below ~50 us) and the same step replayed from a CUDA graph (device time only)."""
import sys

import torch

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
print(f"{'case':14s} {'layout':5s} {'rows':>7s} {'eager us':>9s} {'graph us':>9s}  {'Gvec/s (graph)':>14s}")
for name, shape in (("top B=8", (8, 32, 32, 64)), ("bottom B=8", (8, 64, 64, 64)), ("bottom B=32", (32, 64, 64, 64)),
                    ("bottom B=128", (128, 64, 64, 64))):
    q = vq.Quantize(64, 512).to(dev).train()
    n = shape[0] * shape[1] * shape[2]
    pick = torch.randint(0, 512, (n,), device=dev)
    x = (q.embed.t()[pick] + 0.1 * torch.randn(n, 64, device=dev)).reshape(shape)
    q.cluster_size.data.fill_(n / 512.0); q.embed_avg.data.copy_(q.embed * (n / 512.0))
    for layout in ("dense", "nchw"):
        xx = x if layout == "dense" else x.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
        for _ in range(10):
            q(xx)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(100):
            q(xx)
        b.record()
        torch.cuda.synchronize()
        eager = a.elapsed_time(b) * 10
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):                       # ten chained training steps per replay
                out = q(xx)
        g.replay()
        torch.cuda.synchronize()
        a.record()
        for _ in range(10):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        graph = a.elapsed_time(b) * 10
        print(f"{name:14s} {layout:5s} {n:7d} {eager:9.1f} {graph:9.1f}  {n / graph / 1e3:14.2f}")
