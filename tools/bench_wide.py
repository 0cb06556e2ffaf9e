"""D axis of the cfg-5 sweep (BASELINE.json configs[4]) on the wide tcgen05 engine (csrc/tc_wide_kernel.cuh), one GPU,
N = 524 288 rows, clustered rows (healthy-codebook regime), CUDA events, inputs resident in HBM.
Usage: python tools/bench_wide.py [out.json] [--simt] [--randn]"""
import json
import sys

import torch

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402
from tools.bench_configs import clustered, steady, timeit  # noqa: E402

dev = "cuda:0"


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else None
    N = 128 * 64 * 64
    res = {}
    for D in (128, 256):
        for K in (512, 1024, 2048, 4096, 8192):
            for engine in (("auto", "simt") if ("--simt" in sys.argv and K == 512) else ("auto",)):
                torch.manual_seed(0)
                q = vq.Quantize(D, K, engine=engine).to(dev).train()
                x = (torch.randn(N, D, device=dev) if "--randn" in sys.argv else clustered(q.embed, N, 40)).reshape(128, 64, 64, D)
                steady(q, N)
                steps = 10 if engine == "auto" else 2
                ms_train = timeit(lambda i: q(x), steps, 2)
                q.eval()
                with torch.no_grad():
                    ms_eval = timeit(lambda i: q(x), steps, 2)
                    ms_assign = timeit(lambda i: q.assign(x), steps, 2)
                ws = q._workspace(torch.device(dev), N)
                flagged = int(ws["scratch"][16:20].view(torch.int32).item())
                flops = 2.0 * N * D * K
                res[f"D{D}_K{K}_{engine}"] = {
                    "train_ms": ms_train, "eval_ms": ms_eval, "assign_ms": ms_assign, "flagged_rows_last_call": flagged,
                    "eval_tflops_algorithmic": flops / (ms_eval * 1e-3) / 1e12,
                    "assign_tflops_algorithmic": flops / (ms_assign * 1e-3) / 1e12,
                    "eval_hbm_GBps_algorithmic": N * (8 * D + 8) / (ms_eval * 1e-3) / 1e9,
                    "train_vectors_per_s": N / (ms_train * 1e-3)}
                print(f"D={D} K={K} {engine}: train {ms_train:.3f} ms, eval {ms_eval:.3f} ms ({flops / (ms_eval * 1e-3) / 1e12:.0f} TF/s), "
                      f"assign {ms_assign:.3f} ms ({flops / (ms_assign * 1e-3) / 1e12:.0f} TF/s), flagged {flagged}", flush=True)
                del q, x
                torch.cuda.empty_cache()
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
