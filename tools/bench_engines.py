"""Kernel-alone and whole-step times of every assignment engine at cfg-2 (N = 524 288, D = 64, K = 512), dense rows.
    python tools/bench_engines.py [--dist clustered|randn] [--steps 50]"""
import argparse
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402
from vq_vae_2_pytorch_b200 import _native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dist", default="clustered")
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--engines", default="tcgen05_bf16,tcgen05_tf32,tcgen05")
ap.add_argument("--rows", type=int, default=128 * 64 * 64)
ap.add_argument("--out", default="")
args = ap.parse_args()
dev = torch.device("cuda:0")
lib = _native.load()
D, K, N = 64, 512, args.rows
torch.manual_seed(0)
embed0 = torch.randn(D, K, device=dev)
xs = []
for i in range(3):
    g = torch.Generator(device=dev).manual_seed(1234 + 1000 * i)
    if args.dist == "clustered":
        pick = torch.randint(0, K, (N,), device=dev, generator=g)
        xs.append((embed0.t()[pick] + 0.1 * torch.randn(N, D, device=dev, generator=g)).contiguous())
    else:
        xs.append(torch.randn(N, D, device=dev, generator=g))
res = {}
for name in args.engines.split(","):
    q = vq.Quantize(D, K, engine=name).to(dev).train()

    def reset():
        q.embed.data.copy_(embed0)
        if args.dist == "clustered":
            q.cluster_size.data.fill_(float(N) / K); q.embed_avg.data.copy_(embed0 * (float(N) / K))
        else:
            q.embed_avg.data.copy_(embed0); q.cluster_size.data.zero_()
    reset()
    for i in range(5):
        out = q(xs[i % 3])
    torch.cuda.synchronize()
    reset()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(args.steps):
        q(xs[i % 3])
    b.record()
    torch.cuda.synchronize()
    step_us = a.elapsed_time(b) * 1e3 / args.steps
    # kernel alone
    ws = q._workspace(dev, N)
    reset()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _native.check(lib.vqb200_codebook_prepare(_native.ptr(q.embed), D, K, _native.ptr(ws["image"]), st), "prepare")
    quant = torch.empty(N, D, device=dev); ind = torch.empty(N, dtype=torch.int64, device=dev)
    ws["scratch"][:256].zero_()

    def kern(i):
        _native.check(lib.vqb200_debug_tc_kernel(_native.ptr(xs[i % 3]), N, D, K, _native.ptr(ws["image"]), _native.ptr(quant),
                                                 _native.ptr(ind), _native.ptr(ws["scratch"]), _native.ENGINES[name], st), "kernel")
    for i in range(3):
        kern(i)
    torch.cuda.synchronize()
    flagged = int(ws["scratch"][16:20].view(torch.int32).item()) // 3
    a.record()
    for i in range(args.steps):
        kern(i)
    b.record()
    torch.cuda.synchronize()
    k_us = a.elapsed_time(b) * 1e3 / args.steps
    res[name] = {"step_us": step_us, "kernel_us": k_us, "flagged_rows_per_call": flagged,
                 "hbm_frac_kernel": N * (8 * D + 8) / (k_us * 1e-6) / 1e9 / 6550.1}
    print(f"{name:14s} {args.dist:9s} step {step_us:7.1f} us   kernel {k_us:7.1f} us   flagged/call {flagged}   kernel HBM frac {res[name]['hbm_frac_kernel']:.3f}", flush=True)
if args.out:
    json.dump({"dist": args.dist, "rows": N, "results": res}, open(args.out, "w"), indent=1)
