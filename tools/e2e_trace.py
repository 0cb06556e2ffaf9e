"""Timeline of one vqb200_host_quantize call (CUPTI via torch.profiler): when do the H2D / D2H copies and kernels run?"""
import ctypes as C
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402
from vq_vae_2_pytorch_b200 import _native  # noqa: E402

lib = _native.load()
D, K, N = 64, 512, 128 * 64 * 64
dev = "cuda:0"
torch.manual_seed(0)
q = vq.Quantize(D, K).to(dev).train()
pick = torch.randint(0, K, (N,), device=dev)
x = (q.embed.t()[pick] + 0.1 * torch.randn(N, D, device=dev)).contiguous()
hx = x.cpu().pin_memory()
hq = torch.empty(N, D).pin_memory(); hi = torch.empty(N, dtype=torch.int64).pin_memory(); hd = torch.empty(1).pin_memory()
ctx = C.c_void_p()
_native.check(lib.vqb200_host_ctx_create(N, D, K, C.byref(ctx)), "ctx")


def call():
    _native.check(lib.vqb200_host_quantize(ctx, C.c_void_p(hx.data_ptr()), N, _native.ptr(q.embed), _native.ptr(q.cluster_size),
                                           _native.ptr(q.embed_avg), 0.99, float(1 - 0.99), 1e-5, 1, C.c_void_p(hq.data_ptr()),
                                           C.c_void_p(hi.data_ptr()), C.c_void_p(hd.data_ptr()), 0), "host_quantize")


for _ in range(3):
    call()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    call()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
t0 = min(e.time_range.start for e in evs)
rows = sorted((e.time_range.start - t0, e.time_range.end - t0, e.name[:40]) for e in evs)
print(f"{len(rows)} device activities, span {max(r[1] for r in rows):.0f} us")
for kind in ("Memcpy HtoD", "Memcpy DtoH"):
    r = [x for x in rows if x[2].startswith(kind)]
    busy = sum(b - a for a, b, _ in r)
    print(f"{kind}: {len(r)} copies, first start {r[0][0]:.0f} us, last end {r[-1][1]:.0f} us, busy {busy:.0f} us")
for a, b, n in rows[:40]:
    print(f"{a:8.0f} {b:8.0f} {b - a:7.0f}  {n}")
