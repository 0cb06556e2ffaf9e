"""Trainer-level glue around the quantizer (reference train_vqvae.py:93-118,166-171 and
distributed/distributed.py:75-107).

Once `Quantize` costs ~0.1 ms, what the reference's trainers do AROUND it dominates a data-parallel step:

  * `nn.parallel.DistributedDataParallel(model, ...)` with the default `broadcast_buffers=True` (train_vqvae.py:166-171)
    re-broadcasts every registered buffer from rank 0 before EVERY forward -- for a VQVAE that is the six EMA buffers of
    the two quantizers (2 x 260 KB), one NCCL broadcast per step on the critical path.  It is redundant here: the
    statistics exchange of vqvae.py:58-59 (fused into the EMA kernel) applies the SAME reduced statistics on every rank in
    rank order, so the replicas stay bit-identical by construction (`replicas_identical` proves it on demand).
  * every step ends with `recon_loss.item()` (a host-device synchronisation: the host cannot run ahead any more) and a
    PICKLED `all_gather` of a two-entry dict (distributed.py:75-107: pickle -> byte tensor -> size all_gather -> padded
    all_gather -> unpickle on every rank) just to keep a running mean for the progress bar (train_vqvae.py:93-118).

`ddp_wrap` and `DeferredMetrics` are the drop-in replacements: same numbers, no per-step synchronisation, one 16-byte
all-reduce per logging interval.  PyTorch only provides the plumbing (DDP, NCCL); nothing here touches the hot path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn


def ddp_wrap(model: nn.Module, device: torch.device | int | None = None, **kwargs) -> nn.Module:
    """`DistributedDataParallel` as train_vqvae.py:166-171 builds it, minus the per-forward buffer broadcast.

    Safe with this package's `Quantize`: its EMA buffers are updated from all-reduced statistics in a rank-independent
    summation order, so they never diverge (use `replicas_identical(model)` as an assertion while debugging).  Models with
    BatchNorm running statistics keep PyTorch's default by passing `broadcast_buffers=True` explicitly.
    """
    kwargs.setdefault("broadcast_buffers", False)
    if device is not None and "device_ids" not in kwargs:
        idx = device.index if isinstance(device, torch.device) else int(device)
        kwargs["device_ids"] = [idx]
        kwargs.setdefault("output_device", idx)
    return nn.parallel.DistributedDataParallel(model, **kwargs)


class DeferredMetrics:
    """Running sums kept ON THE DEVICE; replaces the per-step `loss.item()` + pickled `all_gather` of the reference
    trainers (train_vqvae.py:93-118) by one tiny all-reduce and ONE host read per logging interval.

        m = DeferredMetrics(device, ("mse_sum", "mse_n"))
        for img in loader:
            ...
            m.add(mse_sum=recon_loss.detach() * img.shape[0], mse_n=img.shape[0])     # no synchronisation
            if step % 100 == 0:
                tot = m.totals()            # SUM over steps and ranks so far: {'mse_sum': ..., 'mse_n': ...}
                print(tot["mse_sum"] / tot["mse_n"])
    """

    def __init__(self, device, names=("mse_sum", "mse_n")):
        self.names = tuple(names)
        self._index = {n: i for i, n in enumerate(self.names)}
        self._acc = torch.zeros(len(self.names), dtype=torch.float64, device=device)

    def add(self, **values) -> None:
        """Accumulate device tensors or Python numbers; never reads anything back."""
        for name, v in values.items():
            i = self._index[name]
            if isinstance(v, torch.Tensor):
                self._acc[i] += v.detach().to(torch.float64).reshape(())
            else:
                self._acc[i] += float(v)

    def totals(self, group=None) -> dict:
        """SUM over all `add` calls of all ranks (one all-reduce of len(names) doubles, one host read)."""
        t = self._acc.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        vals = t.tolist()
        return dict(zip(self.names, vals))

    def reset(self) -> None:
        self._acc.zero_()


@torch.no_grad()
def replicas_identical(model: nn.Module, group=None) -> bool:
    """True when every registered buffer of every `Quantize` inside `model` is bit-identical on all ranks (debug helper:
    three small all-gathers per quantizer and a host read -- not for the hot loop)."""
    from .quantize import Quantize
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return True
    world = dist.get_world_size(group)
    same = True
    for mod in model.modules():
        if isinstance(mod, Quantize):
            for buf in (mod.embed, mod.cluster_size, mod.embed_avg):
                gathered = [torch.empty_like(buf) for _ in range(world)]
                dist.all_gather(gathered, buf.contiguous(), group=group)
                same &= all(torch.equal(gathered[0], g) for g in gathered[1:])
    return bool(same)
