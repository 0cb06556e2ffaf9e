import torch, time
n = 134217728
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n + 4194304, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n + 4194304, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D 134MB {a:.2f} ms ({n/a/1e6:.1f} GB/s)  D2H 138MB {b:.2f} ms ({(n+4194304)/b/1e6:.1f} GB/s)  both concurrently {c:.2f} ms")
