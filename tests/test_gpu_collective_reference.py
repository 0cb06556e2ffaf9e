"""Row a10 (the collective) against THE REFERENCE ITSELF on two GPUs: the unmodified `Quantize` (oracle/_ref/vqvae.py)
calling the reference's own `dist_fn.all_reduce` (oracle/_ref/distributed/distributed.py:64-72 -> NCCL), one process per
GPU, against the drop-in module on its default data-parallel path (statistics exchange fused into the EMA kernel over peer
memory) -- same per-rank inputs, four chained training steps through tests/ref_harness.compare_step on every rank
(indices exact bar fp64 near-ties, everything else 1e-5 element-wise), replicas bit-identical afterwards.  Also the NCCL
fallback path (VQB200_NO_P2P=1), and the separate in-place exchange kernel that shapes outside the fused kernel take: the
D = 256 deep-fork shape (vqvae_deep.py:252), a sliced codebook (K = 1024) and an odd shape on the SIMT engine.
Needs 2 GPUs (skipped on a 1-GPU box); the per-rank report goes to gpurun_out/parity_report_collective.json."""
import json
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import ref_harness as H
        import vq_vae_2_pytorch_b200 as vq
        from vq_vae_2_pytorch_b200 import replicas_identical
        from oracle import reference_module
        ref = reference_module.load("vqvae")
        H.fp32_reference_backends()
        res = {"paths": {}}
        for name, D, K, shape, no_p2p in (("peer_memory_D64_K512", 64, 512, (16, 64, 64, 64), False),
                                          ("peer_memory_D64_K256_nchw", 64, 256, (8, 32, 32, 64), False),
                                          ("nccl_fallback_D64_K512", 64, 512, (16, 32, 32, 64), True),
                                          ("deep_fork_D256_K512", 256, 512, (4, 32, 32, 256), False),
                                          ("sliced_D64_K1024", 64, 1024, (8, 32, 32, 64), False),
                                          ("odd_D48_K100", 48, 100, (4, 16, 16, 48), False)):
            torch.manual_seed(11)                         # same initial codebook on every rank (DDP would broadcast it)
            r = ref.Quantize(D, K).to(dev).train()
            o = vq.Quantize(D, K).to(dev).train()
            o.load_state_dict(r.state_dict(), strict=True)
            if no_p2p:
                os.environ["VQB200_NO_P2P"] = "1"
            try:
                for step, kind in enumerate(["randn", "randn", "clustered", "clustered"]):
                    x = H.make_inputs(kind, shape, r.embed.detach(), 500 + 100 * step + rank, dev, permuted="nchw" in name)
                    H.compare_step(f"rank{rank}-{name}-step{step}-{kind}", r, o, x)      # (leaves o on the reference's state)
                # two more steps of the candidate alone (no re-synchronisation with the reference): its replicas stay
                # bit-identical by construction (rank-ordered sums / one all-reduce)
                for step in range(2):
                    o(H.make_inputs("clustered", shape, r.embed.detach(), 900 + 10 * step + rank, dev, permuted="nchw" in name))
                assert replicas_identical(o)
            finally:
                os.environ.pop("VQB200_NO_P2P", None)
            took_peer = o._ws[dev]["peer"] is not None
            res["paths"][name] = "peer memory" if took_peer else "nccl"
            # the statistics really were summed over BOTH ranks: cluster_size grew by (1 - decay) * world * rows per step
            res[name + "_cluster_mass"] = float(r.cluster_size.sum())
        res["report"] = H.REPORT
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_two_rank_training_vs_reference_with_its_own_all_reduce():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import reference_module
    if not reference_module.available():
        pytest.skip("reference not staged")
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = {k: dict(v) for k, v in dict(out).items()}
    for rank in (0, 1):
        p = res[rank]["paths"]
        assert p["peer_memory_D64_K512"] == "peer memory" and p["peer_memory_D64_K256_nchw"] == "peer memory", p
        assert p["nccl_fallback_D64_K512"] == "nccl", p
        # shapes outside the fused fold + EMA kernel: the separate in-place exchange kernel (vqb200_stats_exchange_peers)
        assert p["deep_fork_D256_K512"] == p["sliced_D64_K1024"] == p["odd_D48_K100"] == "peer memory", p
        assert len(res[rank]["report"]) == 24
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "parity_report_collective.json"), "w") as f:
        json.dump({"what": "2 x B200, one process per GPU: unmodified reference Quantize + its own dist_fn.all_reduce (NCCL) vs the "
                           "drop-in module, 4 chained training steps per configuration, per rank", **{str(k): v for k, v in res.items()}},
                  f, indent=1)
