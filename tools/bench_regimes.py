"""Eval-mode forward time (assignment + fix-up + outputs, fixed codebook) per engine and input regime at cfg-2 size:
clustered rows (0 rows to re-score), N(0,1) rows against the reference-init codebook (many near-ties), and the same
against a collapsed codebook (dead codes of magnitude ~1e5)."""
import sys

import torch

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402

dev = torch.device("cuda:0")
D, K, N = 64, 512, 128 * 64 * 64
torch.manual_seed(0)
e0 = torch.randn(D, K, device=dev)
collapsed = e0.clone()
collapsed[:, 60:] *= 1.0e5
g = torch.Generator(device=dev).manual_seed(1)
pick = torch.randint(0, K, (N,), device=dev, generator=g)
cases = {"clustered": (e0, (e0.t()[pick] + 0.1 * torch.randn(N, D, device=dev, generator=g)).contiguous()),
         "randn / reference-init codebook": (e0, torch.randn(N, D, device=dev, generator=g)),
         "randn / collapsed codebook": (collapsed, torch.randn(N, D, device=dev, generator=g))}
engines = sys.argv[1].split(",") if len(sys.argv) > 1 else ["tcgen05_bf16", "tcgen05_tf32", "tcgen05", "auto"]
for cname, (emb, x) in cases.items():
    for eng in engines:
        q = vq.Quantize(D, K, engine=eng).to(dev).eval()
        q.embed.data.copy_(emb)
        for _ in range(12):                       # (lets the adaptive policy of engine="auto" settle)
            q(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            q(x)
        b.record()
        torch.cuda.synchronize()
        ws = q._ws[dev]
        flagged = int(ws["scratch"][16:20].view(torch.int32).item())
        print(f"{cname:34s} {eng:13s} eval forward {a.elapsed_time(b) / 20 * 1e3:8.1f} us   rows re-scored exactly: {flagged:6d}"
              + (f"   (policy: {q._filter['mode']})" if eng == "auto" else ""), flush=True)
