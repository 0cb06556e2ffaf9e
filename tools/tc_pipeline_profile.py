"""Per-role cycle breakdown of the tcgen05 kernel (pipeline bubble analysis), via vqb200_debug_tc_profile.
Usage (GPU box):  python tools/tc_pipeline_profile.py [rows]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402
from vq_vae_2_pytorch_b200 import _native  # noqa: E402

SLOTS = ["prod_wait_xe", "mma_wait_af", "mma_wait_te", "mma_total", "conv_wait_xf", "conv_wait_ae", "conv_total",
         "epi0_wait_tf", "epi0_scan", "epi0_total", "epi1_wait_tf", "epi1_scan", "epi1_wait_pf", "epi1_wait_re",
         "epi1_total", "out_wait_rf", "out_total", "kernel", "conv_loop", "conv_tail", "conv_fence"]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128 * 64 * 64
    eng = _native.ENGINES[sys.argv[2]] if len(sys.argv) > 2 else _native.ENGINE_TCGEN05
    lib = _native.load()
    dev = "cuda:0"
    D, K = 64, 512
    torch.manual_seed(0)
    embed = torch.randn(D, K, device=dev)
    pick = torch.randint(0, K, (n,), device=dev)
    x = (embed.t()[pick] + 0.1 * torch.randn(n, D, device=dev)).contiguous()
    image = torch.empty(lib.vqb200_codebook_bytes(D, K), dtype=torch.uint8, device=dev)
    scratch = torch.empty(lib.vqb200_forward_scratch_bytes(n, D, K), dtype=torch.uint8, device=dev)
    quant = torch.empty_like(x)
    ind = torch.empty(n, dtype=torch.int64, device=dev)
    ns = lib.vqb200_tc_profile_slots()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _native.check(lib.vqb200_codebook_prepare(_native.ptr(embed), D, K, _native.ptr(image), st), "prepare")
    for rep in range(3):
        prof = torch.zeros(160, ns, dtype=torch.int64, device=dev)
        _native.check(lib.vqb200_debug_tc_profile(_native.ptr(x), n, D, K, _native.ptr(image), _native.ptr(quant),
                                                  _native.ptr(ind), _native.ptr(scratch), _native.ptr(prof), eng, st), "profile")
        torch.cuda.synchronize()
    p = prof.cpu().double()
    p = p[p[:, SLOTS.index("kernel")] > 0]
    tiles = (n + 127) // 128
    print(f"engine={eng} rows={n} tiles={tiles} ctas={p.shape[0]} tiles/cta~{tiles / p.shape[0]:.1f}")
    print(f"{'slot':16s} {'mean cyc/CTA':>14s} {'per tile':>10s} {'% kernel':>9s}")
    kern = p[:, SLOTS.index("kernel")].mean()
    for i, name in enumerate(SLOTS):
        m = p[:, i].mean().item()
        print(f"{name:16s} {m:14.0f} {m / (tiles / p.shape[0]):10.0f} {100 * m / kern:8.1f}%")


if __name__ == "__main__":
    main()
