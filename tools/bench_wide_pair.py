"""D = 256, K = 512 (the deep fork's quantizer shape, vqvae_deep.py:252,257) at N = 524 288 rows: assign / eval forward / training
step with the wide tensor-core engine; run once with VQB200_TCW_CTA2=0 (two slices of 256 codes, two passes over x) and once with
the default (CTA pair, all 512 codes resident, one pass)."""
import os
import sys

import torch

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402

dev = "cuda:0"
for D, K in [tuple(int(v) for v in a.split('x')) for a in (sys.argv[1:] or ['256x512', '256x1024', '256x2048', '256x8192'])]:
    N = 524288
    torch.manual_seed(0)
    q = vq.Quantize(D, K).to(dev)
    xs = []
    for i in range(3):
        pick = torch.randint(0, K, (N,), device=dev)
        xs.append((q.embed.t()[pick] + 0.1 * torch.randn(N, D, device=dev)).contiguous())
    q.cluster_size.data.fill_(N / K); q.embed_avg.data.copy_(q.embed * (N / K))

    def timed(fn, iters=10):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(iters):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters * 1e3
    q.eval()
    t_assign = timed(lambda i: q.assign(xs[i % 3]))
    t_eval = timed(lambda i: q(xs[i % 3]))
    q.train()
    t_train = timed(lambda i: q(xs[i % 3]))
    print(f"VQB200_TCW_CTA2={os.environ.get('VQB200_TCW_CTA2', '1 (default)')}: D={D} K={K} N={N}: assign {t_assign:.1f} us, eval forward {t_eval:.1f} us, "
          f"training step {t_train:.1f} us", flush=True)
