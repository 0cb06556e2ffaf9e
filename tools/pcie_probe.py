"""PCIe ceiling of the box for the e2e path: pinned H2D alone, D2H alone, and both directions at once (two streams),
with the byte counts of one cfg-2 step (134 MB in, 138 MB out)."""
import torch

dev = "cuda:0"
n_in, n_out = 134217728, 138412036
hin = torch.empty(n_in, dtype=torch.uint8).pin_memory(); hout = torch.empty(n_out, dtype=torch.uint8).pin_memory()
din = torch.empty(n_in, dtype=torch.uint8, device=dev); dout = torch.empty(n_out, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=10):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_event(a); s2.wait_event(a)
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
    e1, e2 = torch.cuda.Event(), torch.cuda.Event()
    e1.record(s1); e2.record(s2)
    torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for _ in range(2):
    run(True, True, 2)
t = run(True, False); print(f"H2D alone : {t:.3f} ms  {n_in / t / 1e6:.1f} GB/s")
t = run(False, True); print(f"D2H alone : {t:.3f} ms  {n_out / t / 1e6:.1f} GB/s")
t = run(True, True); print(f"both      : {t:.3f} ms  {n_in / t / 1e6:.1f} + {n_out / t / 1e6:.1f} GB/s")
