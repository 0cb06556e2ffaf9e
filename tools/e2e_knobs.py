"""e2e (vqb200_host_quantize, cfg-2, pinned host buffers) under the host-path knobs: wall-clock ms per call, median of 15."""
import ctypes as C
import os
import sys
import time

import torch

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
q.cluster_size.data.fill_(N / K); q.embed_avg.data.copy_(q.embed * (N / K))
hx = x.cpu().pin_memory()
hq = torch.empty(N, D).pin_memory(); hi = torch.empty(N, dtype=torch.int64).pin_memory(); hd = torch.empty(1).pin_memory()
for knobs in sys.argv[1:] or [""]:
    for kv in filter(None, knobs.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    ctx = C.c_void_p()
    _native.check(lib.vqb200_host_ctx_create(N, D, K, C.byref(ctx)), "ctx")
    ts = []
    for i in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _native.check(lib.vqb200_host_quantize(ctx, C.c_void_p(hx.data_ptr()), N, _native.ptr(q.embed), _native.ptr(q.cluster_size),
                                               _native.ptr(q.embed_avg), 0.99, float(1 - 0.99), 1e-5, 1, C.c_void_p(hq.data_ptr()),
                                               C.c_void_p(hi.data_ptr()), C.c_void_p(hd.data_ptr()), 0), "host_quantize")
        ts.append((time.perf_counter() - t0) * 1e3)
    lib.vqb200_host_ctx_destroy(ctx)
    ts = sorted(ts[5:])
    print(f"{knobs or 'default':50s} median {ts[len(ts) // 2]:.3f} ms  min {ts[0]:.3f}  max {ts[-1]:.3f}", flush=True)
    for kv in filter(None, knobs.split(",")):
        os.environ.pop(kv.split("=")[0], None)
