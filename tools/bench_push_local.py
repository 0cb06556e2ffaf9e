"""Single GPU: the multi-rank code path with world = 1 (fold kernel pushes to a local slot, EMA kernel waits on the local
flag) against the single-rank fused step -- isolates what the exchange path costs without any NVLink traffic."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402
from vq_vae_2_pytorch_b200 import _native  # noqa: E402

dev = torch.device("cuda:0")
lib = _native.load()
D, K, N = 64, 512, 128 * 64 * 64
torch.manual_seed(0)
q = vq.Quantize(D, K).to(dev).train()
e0 = q.embed.clone()
xs = []
for i in range(3):
    g = torch.Generator(device=dev).manual_seed(1234 + 1000 * i)
    pick = torch.randint(0, K, (N,), device=dev, generator=g)
    xs.append((e0.t()[pick] + 0.1 * torch.randn(N, D, device=dev, generator=g)).contiguous())
ws = q._workspace(dev, N)
n = lib.vqb200_stats_bytes(D, K) // 4
n_al = (n + 63) // 64 * 64
buf = torch.zeros(2 * n_al + 128, device=dev)
quant = torch.empty(N, D, device=dev); ind = torch.empty(N, dtype=torch.int64, device=dev); diff = torch.empty((), device=dev)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
eng = _native.ENGINE_TCGEN05_BF16


def reset():
    q.embed.data.copy_(e0); q.cluster_size.data.fill_(N / K); q.embed_avg.data.copy_(e0 * (N / K))


def fused(i, step):
    _native.check(lib.vqb200_quantize_step(xs[i % 3].data_ptr(), N, D, K, N, 0, D, 1, q.embed.data_ptr(), q.cluster_size.data_ptr(),
                                           q.embed_avg.data_ptr(), ws["image"].data_ptr(), quant.data_ptr(), ind.data_ptr(),
                                           diff.data_ptr(), ws["stats"].data_ptr(), ws["scratch"].data_ptr(), None, eng, 1, 0.99,
                                           float(1 - 0.99), 1e-5, st), "step")


def split(i, step):
    _native.check(lib.vqb200_quantize_step(xs[i % 3].data_ptr(), N, D, K, N, 0, D, 1, q.embed.data_ptr(), q.cluster_size.data_ptr(),
                                           q.embed_avg.data_ptr(), ws["image"].data_ptr(), quant.data_ptr(), ind.data_ptr(),
                                           diff.data_ptr(), ws["stats"].data_ptr(), ws["scratch"].data_ptr(), None, eng, 0, 0.99,
                                           float(1 - 0.99), 1e-5, st), "step")
    _native.check(lib.vqb200_ema_update(ws["stats"].data_ptr(), q.cluster_size.data_ptr(), q.embed_avg.data_ptr(), q.embed.data_ptr(),
                                        D, K, 0.99, float(1 - 0.99), 1e-5, None, st), "ema")


nw = K * (D + 1)
nw_al = (nw + 63) // 64 * 64
pbuf = torch.zeros(2 * 2 * nw_al + 64, device=dev)


def peers_w1(i, step):
    slot = [pbuf.data_ptr() + 4 * par * 2 * nw_al for par in (0, 1)]
    err = pbuf.data_ptr() + 4 * (4 * nw_al)
    dst = (C.c_void_p * 2)(*slot); rc = (C.c_void_p * 2)(*slot)
    _native.check(lib.vqb200_quantize_step_peers(xs[i % 3].data_ptr(), N, D, K, N, 0, D, 1, q.embed.data_ptr(), q.cluster_size.data_ptr(),
                                                 q.embed_avg.data_ptr(), ws["image"].data_ptr(), quant.data_ptr(), ind.data_ptr(),
                                                 diff.data_ptr(), ws["scratch"].data_ptr(), None, eng, 0.99, float(1 - 0.99), 1e-5,
                                                 dst, rc, C.c_void_p(err), C.c_void_p(err + 128), 0, 1, st), "peers")


step_no = 0
for name, fn in (("fused single-rank step (fold inside the EMA kernel)", fused), ("forward + k_stats_fold + separate local EMA", split),
                 ("one-call multi-rank step, world 1 (fold + push + flags + EMA in one kernel)", peers_w1)):
    reset()
    for i in range(6):
        step_no += 1
        fn(i, step_no)
    torch.cuda.synchronize()
    reset()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(50):
        step_no += 1
        fn(i, step_no)
    b.record()
    torch.cuda.synchronize()
    print(f"{name:78s} {a.elapsed_time(b) / 50 * 1e3:7.1f} us/step", flush=True)
