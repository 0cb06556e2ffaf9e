"""Small end-to-end run for compute-sanitizer (memcheck): dense, NCHW-physical and sliced-codebook forwards + backward."""
import sys

import torch

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
for K in (512, 1024):
    q = vq.Quantize(64, K).to(dev).train()
    for shape, nchw in (((2, 16, 16, 64), False), ((2, 64, 16, 16), True), ((300, 64), False)):
        x = torch.randn(*shape, device=dev)
        if nchw:
            x = x.permute(0, 2, 3, 1)
        x.requires_grad_(True)
        for _ in range(2):
            quant, diff, ind = q(x)
            (quant.sum() + 0.25 * diff).backward()
    q.eval()
    q(torch.randn(1000, 64, device=dev))
    q.assign(torch.randn(129, 64, device=dev))
torch.cuda.synchronize()
print("sanitize_small: done")
