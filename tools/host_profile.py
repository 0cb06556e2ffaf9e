"""Where the ~49 us of an eager training forward go on the host: cProfile over 3000 small calls (B = 8 bottom shape)."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402

dev = "cuda:0"
q = vq.Quantize(64, 512).to(dev).train()
x = torch.randn(8, 64, 64, 64, device=dev)
for _ in range(50):
    q(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3000):
    q(x)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host issue time per call: {(t1 - t0) / 3000 * 1e6:.1f} us")
pr = cProfile.Profile()
pr.enable()
for _ in range(3000):
    q(x)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(14)
