"""Collective shim of the quantizer path.

Mirrors the two helpers of the reference that the hot path touches
(/root/reference/distributed/distributed.py:54-61 `get_world_size`, :64-72 `all_reduce`):
world-size-1 (or an uninitialised process group) short-circuits, otherwise an in-place SUM
all-reduce over the default group.  The reference issues TWO all-reduces per quantizer forward
(vqvae.py:58-59: counts [K], then sums [D,K]); here both travel in ONE packed fp32 buffer
`[K*D sums | K counts]` (`packed_stats_numel`), so the step pays one NCCL launch latency.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def get_world_size(group=None) -> int:
    if not dist.is_available() or not dist.is_initialized():
        return 1
    return dist.get_world_size(group)


def all_reduce(tensor: torch.Tensor, op=None, group=None) -> torch.Tensor:
    """In-place SUM all-reduce; no-op for a single process (reference distributed.py:64-72)."""
    if get_world_size(group) == 1:
        return tensor
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM if op is None else op, group=group)
    return tensor


def packed_stats_numel(dim: int, n_embed: int) -> int:
    """Number of fp32 elements of the packed statistics buffer: n_embed*dim sums + n_embed counts."""
    return n_embed * (dim + 1)


def split_packed_stats(stats: torch.Tensor, dim: int, n_embed: int):
    """Views into the packed buffer: (embed_sum as [n_embed, dim] code-major, counts [n_embed])."""
    flat = stats.view(-1)
    return flat[: n_embed * dim].view(n_embed, dim), flat[n_embed * dim: n_embed * (dim + 1)]
