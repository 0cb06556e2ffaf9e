"""One eval forward of the wide tensor-core engine (for ncu captures).  Usage: python tools/wide_once.py D K [n_calls]"""
import sys

import torch

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402
from tools.bench_configs import clustered  # noqa: E402

D, K = int(sys.argv[1]), int(sys.argv[2])
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(0)
q = vq.Quantize(D, K).to("cuda:0").eval()
x = clustered(q.embed, 128 * 64 * 64, 40).reshape(128, 64, 64, D)
with torch.no_grad():
    for _ in range(calls):
        quant, diff, ind = q(x)
torch.cuda.synchronize()
print("diff", float(diff), "codes used", int(ind.unique().numel()))
