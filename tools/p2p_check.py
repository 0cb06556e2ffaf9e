"""Multi-GPU check of the fused peer-to-peer all-reduce + EMA kernel against the NCCL all-reduce path.
Run: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/p2p_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    a = vq.Quantize(64, 512).to(dev).train()            # peer-memory path
    b = vq.Quantize(64, 512).to(dev).train()            # NCCL path
    b.load_state_dict(a.state_dict())
    n = 128 * 40 + 17
    b._workspace(dev, n)["peer"] = None
    embed0 = a.embed.clone()
    ok = True
    for step in range(6):
        g = torch.Generator(device=dev).manual_seed(100 * step + rank)
        pick = torch.randint(0, 512, (n,), device=dev, generator=g)
        x = embed0.t()[pick] + 0.3 * torch.randn(n, 64, device=dev, generator=g)
        qa, da, ia = a(x)
        qb, db, ib = b(x)
        same = torch.equal(ia, ib) and torch.allclose(qa, qb, rtol=1e-6, atol=1e-6)
        if not same:
            print(f"[rank {rank}] step {step}: outputs differ ({int((ia != ib).sum())} indices)", flush=True)
        tol = 1e-6                              # vs NCCL: summation order / contraction may differ in the last bit
        for name in ("embed", "cluster_size", "embed_avg"):
            ta, tb = getattr(a, name), getattr(b, name)
            err = float((ta - tb).abs().max() / tb.abs().max())
            if err > tol:
                print(f"[rank {rank}] step {step} {name}: rel err vs NCCL path {err:.3e}", flush=True)
            same = same and err <= tol
            # replicas must be bit-identical across ranks
            ref = ta.clone()
            dist.broadcast(ref, 0)
            if not torch.equal(ref, ta):
                print(f"[rank {rank}] step {step} {name}: replica differs from rank 0 by {float((ref - ta).abs().max()):.3e}", flush=True)
            same = same and torch.equal(ref, ta)
        ok = ok and same
    used_p2p = a._ws[dev].get("peer") is not None
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"p2p_check world={world} peer_memory_path={used_p2p} "
              f"{'(fallback reason: ' + getattr(a, '_peer_error', '?') + ')' if not used_p2p else ''} "
              f"result={'OK' if int(flag.item()) else 'MISMATCH'}")
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
