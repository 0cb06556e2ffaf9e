"""Generate golden vectors by EXECUTING the reference `Quantize` (read-only mount).

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports `/root/reference/vqvae.py` unmodified (torch CPU, fp32, default
matmul precision), drives `Quantize` with seeded inputs and stores inputs,
initial buffers, per-step outputs and per-step buffers in `tests/golden/*.npz`.
Nothing from the reference is copied into this repository; only its outputs.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("VQ_REFERENCE_DIR", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from vqvae import Quantize  # noqa: E402  (the reference module itself)

OUT = os.path.dirname(os.path.abspath(__file__))


def run_case(name, dim, n_embed, xs, *, train=True, embed=None, grads=None, decay=0.99, eps=1e-5):
    torch.manual_seed(0)
    q = Quantize(dim, n_embed, decay=decay, eps=eps)
    if embed is not None:
        q.embed.data.copy_(embed)
        q.embed_avg.data.copy_(embed)
    q.train(train)
    rec = {"dim": dim, "n_embed": n_embed, "decay": decay, "eps": eps, "train": int(train),
           "steps": len(xs), "embed0": q.embed.numpy().copy(),
           "cluster_size0": q.cluster_size.numpy().copy(), "embed_avg0": q.embed_avg.numpy().copy()}
    for s, x in enumerate(xs):
        x = x.clone().requires_grad_(grads is not None)
        quant, diff, ind = q(x)
        rec[f"x{s}"] = x.detach().numpy().copy()
        rec[f"x{s}_strides"] = np.array(x.stride(), dtype=np.int64)
        rec[f"quantize{s}"] = quant.detach().numpy().copy()
        rec[f"quantize{s}_strides"] = np.array(quant.stride(), dtype=np.int64)
        rec[f"diff{s}"] = diff.detach().numpy().copy()
        rec[f"ind{s}"] = ind.numpy().copy()
        rec[f"embed{s + 1}"] = q.embed.numpy().copy()
        rec[f"cluster_size{s + 1}"] = q.cluster_size.numpy().copy()
        rec[f"embed_avg{s + 1}"] = q.embed_avg.numpy().copy()
        if grads is not None:
            gq, gd = grads[s]
            (quant * gq).sum().add(diff * gd).backward()
            rec[f"gq{s}"] = gq.numpy().copy()
            rec[f"gd{s}"] = np.float32(gd)
            rec[f"xgrad{s}"] = x.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, {k: getattr(v, "shape", v) for k, v in rec.items() if k.startswith(("x0", "ind0"))})


def gen(seed):
    return torch.Generator().manual_seed(seed)


def main():
    D, K = 64, 512
    # 1. reference-init regime, 3 EMA steps (dead codes blow up to ~1e5 after step 1)
    run_case("randn_train3", D, K, [torch.randn(2, 16, 16, D, generator=gen(1234 + 1000 * s)) for s in range(3)])
    # 2. what VQVAE.encode really passes: a permute(0,2,3,1) view of an NCHW tensor (vqvae.py:227)
    run_case("permuted_eval", D, K, [torch.randn(2, D, 8, 8, generator=gen(7)).permute(0, 2, 3, 1)], train=False)
    run_case("permuted_train", D, K, [torch.randn(2, D, 8, 8, generator=gen(8)).permute(0, 2, 3, 1)], train=True)
    # 3. healthy-codebook regime: inputs clustered around codes
    torch.manual_seed(0)
    e0 = torch.randn(D, K)
    cl = []
    for s in range(2):
        g = gen(4321 + s)
        pick = torch.randint(0, K, (1024,), generator=g)
        cl.append((e0[:, pick].t() + 0.1 * torch.randn(1024, D, generator=g)).reshape(4, 16, 16, D).contiguous())
    run_case("clustered_train2", D, K, cl)
    # 4. adversarial ties: duplicate columns, inputs equal to a code, exact midpoints
    g = gen(99)
    e = torch.randn(D, 64, generator=g)
    e[:, 5] = e[:, 2]
    e[:, 63] = e[:, 2]
    e[:, 40] = e[:, 17]
    x = torch.randn(96, D, generator=g)
    x[0:8] = e[:, 2]                      # equal to a duplicated code -> lowest index 2
    x[8:16] = e[:, 17]
    x[16:24] = e[:, [3, 9, 11, 30, 31, 33, 60, 62]].t()
    x[24:32] = 0.5 * (e[:, 20] + e[:, 21])  # midpoint (fp32 rounding decides; tolerated as near-tie)
    x[32:40] = 0.0
    run_case("ties_eval", D, 64, [x], train=False, embed=e)
    run_case("ties_train", D, 64, [x], train=True, embed=e)
    # 5. deep variant shapes: dim 256 (vqvae_deep.py:252,257), 2-D input
    run_case("deep_d256_train2", 256, 128, [torch.randn(2, 9, 6, 256, generator=gen(55 + s)) for s in range(2)])
    run_case("rows2d_eval", D, K, [torch.randn(5, D, generator=gen(3))], train=False)
    # 6. implied backward (vqvae.py:72-73)
    xs = [torch.randn(3, 8, 8, D, generator=gen(2024 + s)) for s in range(2)]
    gr = [(torch.randn(3, 8, 8, D, generator=gen(11 + s)), 0.25 + s) for s in range(2)]
    run_case("backward_train2", D, K, xs, grads=gr)
    # 7. odd sizes: N not a multiple of any tile, small K, non-default decay/eps
    run_case("ragged_train2", 32, 40, [torch.randn(131, 32, generator=gen(77 + s)) * 3.0 for s in range(2)],
             decay=0.9, eps=1e-3)


if __name__ == "__main__":
    main()
