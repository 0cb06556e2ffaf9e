"""Time of the statistics kernels alone (forward with statistics minus forward without), cfg-2, dense and NCHW-physical rows,
uniform and skewed code usage.   python tools/bench_stats.py"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402
from vq_vae_2_pytorch_b200 import _native, row_layout  # noqa: E402

import argparse
ap = argparse.ArgumentParser()
ap.add_argument("--live", default="512,40,3")
ap.add_argument("--layouts", default="dense,nchw")
ap.add_argument("--iters", type=int, default=30)
args = ap.parse_args()
dev = torch.device("cuda:0")
lib = _native.load()
D, K, B, H, W = 64, 512, 128, 64, 64
N = B * H * W
torch.manual_seed(0)
q = vq.Quantize(D, K).to(dev).train()
ws = q._workspace(dev, N)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
_native.check(lib.vqb200_codebook_prepare(_native.ptr(q.embed), D, K, _native.ptr(ws["image"]), st), "prepare")
for live in [int(v) for v in args.live.split(',')]:
    codes = torch.randperm(K, device=dev)[:live]
    xs = []
    for i in range(3):
        pick = codes[torch.randint(0, live, (N,), device=dev)]
        xs.append((q.embed.t()[pick] + 0.1 * torch.randn(N, D, device=dev)).reshape(B, H, W, D).contiguous())
    for layout in args.layouts.split(","):
        xl = xs if layout == "dense" else [x.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1) for x in xs]
        quant = torch.empty_strided(xl[0].shape, xl[0].stride(), device=dev)
        ind = torch.empty(B, H, W, dtype=torch.int64, device=dev); diff = torch.empty((), device=dev)
        n, rpi, img, row, col = row_layout(xl[0])

        def fwd(i, stats):
            _native.check(lib.vqb200_quantize_forward(_native.ptr(xl[i % 3]), n, D, K, rpi, img, row, col, _native.ptr(ws["image"]),
                                                      _native.ptr(quant), _native.ptr(ind), _native.ptr(diff),
                                                      _native.ptr(ws["stats"]) if stats else None, _native.ptr(ws["scratch"]),
                                                      _native.ENGINE_TCGEN05_BF16, st), "fwd")
        t = {}
        for stats in (False, True):
            for i in range(3):
                fwd(i, stats)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(args.iters):
                fwd(i, stats)
            b.record()
            torch.cuda.synchronize()
            t[stats] = a.elapsed_time(b) / args.iters * 1e3
        print(f"live codes {live:3d} {layout:5s}: forward {t[False]:6.1f} us, with statistics {t[True]:6.1f} us -> statistics kernels {t[True] - t[False]:5.1f} us", flush=True)
