"""Index egress for code extraction: the B200-side replacement of the per-batch `id.detach().cpu().numpy()` of
extract_code.py:23-33.

The reference copies the int64 `id_t` / `id_b` of every batch to pageable host memory synchronously.  Here the codes are
narrowed on the device (`vqb200_pack_indices`: uint16 when n_embed <= 65536, else int32 -- 4x / 2x fewer bytes over PCIe),
copied with an asynchronous D2H into a ring of pinned host buffers on a side stream, and handed out in push order once
their copy has finished, so the encoder of batch i+1 overlaps the egress of batch i.  What a consumer stores is unchanged:
`pop()` returns numpy arrays shaped like the index tensors, int64 by default (what `LMDBDataset` feeds to
`torch.from_numpy`, dataset.py:45-51), or the narrow dtype with `widen=False`.
"""
from __future__ import annotations

import collections

import numpy as np
import torch

from . import _native


class CodeEgress:
    def __init__(self, n_embed: int, depth: int = 2, widen: bool = True):
        if n_embed <= 0:
            raise ValueError("n_embed must be positive")
        self.n_embed, self.depth, self.widen = int(n_embed), max(1, int(depth)), bool(widen)
        self.code_bytes = 2 if n_embed <= 65536 else 4
        self._np_dtype = np.uint16 if self.code_bytes == 2 else np.int32
        self._torch_dtype = torch.int16 if self.code_bytes == 2 else torch.int32   # int16 storage, read back as uint16
        self._slots = []                       # free ring slots: dict(dev, host, status_dev, status_host, event, capacity)
        self._pending = collections.deque()    # (slot, [(offset, shape)], tag)
        self._stream = None

    def _slot(self, device, n_codes):
        for i, s in enumerate(self._slots):
            if s["capacity"] >= n_codes and s["dev"].device == device:
                return self._slots.pop(i)
        cap = max(n_codes, 1)
        return {"dev": torch.empty(cap, dtype=self._torch_dtype, device=device),
                "host": torch.empty(cap, dtype=self._torch_dtype).pin_memory(),
                "status_dev": torch.zeros(4, dtype=torch.int32, device=device),
                "status_host": torch.zeros(4, dtype=torch.int32).pin_memory(),
                "event": torch.cuda.Event(), "capacity": cap}

    def push(self, *index_tensors, tag=None):
        """Queue the egress of one batch of index tensors (int64 CUDA tensors, e.g. id_t and id_b).  Returns immediately;
        blocks only when `depth` batches are already in flight (then the oldest is completed first and kept for pop())."""
        if not index_tensors:
            raise ValueError("push() needs at least one index tensor")
        dev = index_tensors[0].device
        for t in index_tensors:
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.int64 and t.device == dev):
                raise RuntimeError("CodeEgress.push expects int64 CUDA tensors on one device (there is no CPU path)")
        if len(self._pending) >= self.depth:           # back-pressure: at most `depth` copies in flight
            self._pending[len(self._pending) - self.depth][0]["event"].synchronize()
        lib = _native.load()
        if self._stream is None or self._stream.device != dev:
            self._stream = torch.cuda.Stream(device=dev)
        pieces, total = [], 0
        for t in index_tensors:
            pieces.append((total, tuple(t.shape), t.numel()))
            total += (t.numel() + 7) // 8 * 8          # 16-byte aligned pieces
        slot = self._slot(dev, total)
        cur = torch.cuda.current_stream(dev)
        self._stream.wait_stream(cur)                  # the indices are produced on the caller's stream
        with torch.cuda.stream(self._stream):
            st = _native.C.c_void_p(self._stream.cuda_stream)
            for k, (t, (off, _, n)) in enumerate(zip(index_tensors, pieces)):
                tc = t if t.is_contiguous() else t.contiguous()
                tc.record_stream(self._stream)
                _native.check(lib.vqb200_pack_indices(tc.data_ptr(), n, self.n_embed, self.code_bytes,
                                                      slot["dev"].data_ptr() + off * self.code_bytes,
                                                      slot["status_dev"].data_ptr() + 4 * min(k, 3), st), "vqb200_pack_indices")
            slot["host"][:total].copy_(slot["dev"][:total], non_blocking=True)
            slot["status_host"].copy_(slot["status_dev"], non_blocking=True)
            slot["event"].record(self._stream)
        self._pending.append((slot, pieces, tag))
        return total * self.code_bytes                 # D2H bytes queued for this batch

    def __len__(self):
        return len(self._pending)

    def pop(self):
        """Oldest batch: (tag, [numpy arrays shaped like the pushed tensors]).  Waits for its copy only."""
        if not self._pending:
            raise IndexError("CodeEgress.pop: nothing in flight")
        slot, pieces, tag = self._pending.popleft()
        slot["event"].synchronize()
        if int(slot["status_host"].max()) != 0:
            self._slots.append(slot)
            raise RuntimeError("CodeEgress: an index outside [0, n_embed) was pushed")
        flat = slot["host"].numpy().view(self._np_dtype)
        out = []
        for off, shape, n in pieces:
            a = flat[off:off + n]
            out.append((a.astype(np.int64) if self.widen else a.copy()).reshape(shape))
        self._slots.append(slot)
        return tag, out

    def drain(self):
        while self._pending:
            yield self.pop()


def unpack_codes(codes: torch.Tensor) -> torch.Tensor:
    """uint16 / int32 CUDA codes -> int64 indices for `Quantize.embed_code` / `VQVAE.decode_code` (vqvae.py:251-259)."""
    if not codes.is_cuda or codes.dtype not in (torch.uint16, torch.int16, torch.int32):
        raise RuntimeError("unpack_codes expects a uint16 / int32 CUDA tensor")
    c = codes.contiguous()
    out = torch.empty(c.shape, dtype=torch.int64, device=c.device)
    st = _native.C.c_void_p(torch.cuda.current_stream(c.device).cuda_stream)
    _native.check(_native.load().vqb200_unpack_indices(c.data_ptr(), c.numel(), c.element_size(), out.data_ptr(), st),
                  "vqb200_unpack_indices")
    return out
