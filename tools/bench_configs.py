"""Secondary configurations of BASELINE.json / SURVEY 8(d) on one GPU (CUDA events, inputs resident in HBM):
  cfg-2n  bottom quantizer, NCHW-physical input (what VQVAE.encode passes)          train fwd+EMA
  cfg-3   top [B,32,32,64] + bottom [B,64,64,64], B = 256, train fwd+EMA (one optimiser step's worth = 2 forwards)
  cfg-4   inference, B = 1024: top 1 048 576 + bottom 4 194 304 rows, argmin only (Quantize.assign) and full eval forward
  cfg-5   K x D sweep at N = 524 288 (shapes outside D = 64, K <= 512 run on the exact SIMT engine)
Usage: python tools/bench_configs.py [out.json] [--sweep]"""
import json
import sys

import torch

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402

dev = "cuda:0"


def clustered(embed, n, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    K = embed.shape[1]
    pick = torch.randint(0, K, (n,), device=dev, generator=g)
    return embed.t()[pick] + 0.1 * torch.randn(n, embed.shape[0], device=dev, generator=g)


def steady(q, n_rows):
    e = q.embed.clone()
    q.cluster_size.data.fill_(float(n_rows) / q.n_embed)
    q.embed_avg.data.copy_(e * (float(n_rows) / q.n_embed))


def timeit(fn, steps, warmup=5):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else None
    res = {}
    torch.manual_seed(0)
    D, K = 64, 512
    # ---- cfg-3
    B = 256
    qt, qb = vq.Quantize(D, K).to(dev).train(), vq.Quantize(D, K).to(dev).train()
    nt, nb = B * 32 * 32, B * 64 * 64
    xt = [clustered(qt.embed, nt, 10 + i).reshape(B, 32, 32, D) for i in range(2)]
    xb = [clustered(qb.embed, nb, 20 + i).reshape(B, 64, 64, D) for i in range(2)]
    steady(qt, nt); steady(qb, nb)
    ms = timeit(lambda i: (qt(xt[i % 2]), qb(xb[i % 2])), 30)
    res["cfg3_top_plus_bottom_B256_train"] = {"ms_per_step": ms, "vectors_per_s": (nt + nb) / (ms * 1e-3), "rows": nt + nb}
    del xt, xb
    # ---- cfg-4
    B = 1024
    qt.eval(); qb.eval()
    nt, nb = B * 32 * 32, B * 64 * 64
    xt = clustered(qt.embed, nt, 30).reshape(B, 32, 32, D)
    xb = clustered(qb.embed, nb, 31).reshape(B, 64, 64, D)
    ms = timeit(lambda i: (qt.assign(xt), qb.assign(xb)), 10, 3)
    res["cfg4_inference_B1024_argmin_only"] = {"ms_per_step": ms, "vectors_per_s": (nt + nb) / (ms * 1e-3), "rows": nt + nb,
                                               "hbm_GBps_algorithmic": (nt + nb) * (4 * D + 8) / (ms * 1e-3) / 1e9}
    with torch.no_grad():
        ms = timeit(lambda i: (qt(xt), qb(xb)), 10, 3)
    res["cfg4_inference_B1024_full_eval_forward"] = {"ms_per_step": ms, "vectors_per_s": (nt + nb) / (ms * 1e-3), "rows": nt + nb,
                                                      "hbm_GBps_algorithmic": (nt + nb) * (8 * D + 8) / (ms * 1e-3) / 1e9}
    del xt, xb
    # ---- cfg-5 sweep
    if "--sweep" in sys.argv:
        N = 128 * 64 * 64
        for Dd in ((64,) if "--d64" in sys.argv else (64, 128, 256)):
            for Kk in (512, 1024, 2048, 4096, 8192):
                torch.manual_seed(0)
                q = vq.Quantize(Dd, Kk).to(dev).train()
                x = clustered(q.embed, N, 40).reshape(128, 64, 64, Dd)
                steady(q, N)
                steps = 20 if Dd == 64 else 3
                ms = timeit(lambda i: q(x), steps, 2)
                res[f"cfg5_D{Dd}_K{Kk}_train"] = {"ms_per_step": ms, "vectors_per_s": N / (ms * 1e-3),
                                                   "tflops_algorithmic": 2.0 * N * Dd * Kk / (ms * 1e-3) / 1e12,
                                                   "engine": ("tcgen05" if Kk <= 512 else f"tcgen05, {Kk // 512} codebook slices") if Dd == 64 else "simt (exact fp32)"}
                del q, x
                torch.cuda.empty_cache()
    print(json.dumps(res, indent=1))
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
