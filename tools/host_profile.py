"""cProfile of the module's host path (tiny input: the GPU is never the limiter)."""
import cProfile
import pstats
import sys

import torch

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402

q = vq.Quantize(64, 512).cuda().train()
x = torch.randn(2, 16, 16, 64, device="cuda")
for _ in range(50):
    q(x)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(2000):
    q(x)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)
