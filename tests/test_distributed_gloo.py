"""world_size-2 gloo test (CPU) of the N>1 host logic: the packed [K*D | K] statistics buffer
all-reduced ONCE equals the reference's two separate all-reduces (vqvae.py:58-59), and applying the
reduced statistics reproduces the oracle run on the concatenated batch."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vq_vae_2_pytorch_b200 import distributed as dist_fn
from oracle.quantize_oracle import QuantizeOracle, code_statistics, distances_f32, ema_update, nearest_code


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        D, K = 16, 24
        embed = np.random.default_rng(0).standard_normal((D, K)).astype(np.float32)
        x = np.random.default_rng(100 + rank).standard_normal((50 + 7 * rank, D)).astype(np.float32)
        ind = nearest_code(distances_f32(x, embed))
        counts, sums = code_statistics(x, ind, K)                       # sums is [D, K]
        # packed layout of the C ABI: K*D code-major sums, then K counts (+ spare scalars not reduced)
        packed = torch.zeros(dist_fn.packed_stats_numel(D, K) + 4)
        s_view, c_view = dist_fn.split_packed_stats(packed[: K * (D + 1)], D, K)
        s_view.copy_(torch.from_numpy(np.ascontiguousarray(sums.T)))
        c_view.copy_(torch.from_numpy(counts))
        packed[-4:] = 123.0                                              # must not travel
        assert dist_fn.get_world_size() == world
        dist_fn.all_reduce(packed[: K * (D + 1)])
        # the reference's way: two separate reductions
        c2 = torch.from_numpy(counts.copy())
        s2 = torch.from_numpy(sums.copy())
        dist.all_reduce(c2)
        dist.all_reduce(s2)
        assert torch.equal(c_view, c2)
        assert torch.allclose(s_view.t(), s2, rtol=0, atol=0)
        assert float(packed[-1]) == 123.0
        cs, ea, e = ema_update(np.zeros(K, np.float32), embed.copy(), c_view.numpy(), s_view.numpy().T.copy(),
                               0.99, 1e-5, K)
        out[rank] = (cs, ea, e)
    finally:
        dist.destroy_process_group()


def test_packed_allreduce_equals_two_reductions_and_global_batch():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    # replicas stay bit-identical (same reduced statistics on every rank)
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    # and equal the single-process oracle on the concatenated batch
    D, K = 16, 24
    embed = np.random.default_rng(0).standard_normal((D, K)).astype(np.float32)
    xs = [np.random.default_rng(100 + r).standard_normal((50 + 7 * r, D)).astype(np.float32) for r in range(world)]
    o = QuantizeOracle(D, K, embed=embed)
    o.forward(np.concatenate(xs, 0))
    assert np.allclose(out[0][0], o.cluster_size, rtol=1e-6, atol=1e-7)
    assert np.allclose(out[0][1], o.embed_avg, rtol=1e-5, atol=1e-6)
    assert np.allclose(out[0][2], o.embed, rtol=1e-4, atol=1e-5)


def test_world_size_one_short_circuits():
    t = torch.ones(4)
    assert dist_fn.get_world_size() == 1
    assert dist_fn.all_reduce(t) is t and torch.equal(t, torch.ones(4))
