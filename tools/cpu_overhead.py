"""Host-side cost of one Quantize.forward call (tiny input so the GPU is never the limiter)."""
import sys, time, torch
sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq
q = vq.Quantize(64, 512).cuda().train()
x = torch.randn(256, 64, device="cuda")
for _ in range(20): q(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 300
for _ in range(n): q(x)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"train-mode forward: host issue time {1e6*(t1-t0)/n:.1f} us/call, incl. drain {1e6*(t2-t0)/n:.1f} us/call")
q.eval()
for _ in range(20): q(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n): q(x)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"eval-mode forward : host issue time {1e6*(t1-t0)/n:.1f} us/call")
