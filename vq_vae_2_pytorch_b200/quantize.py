"""`Quantize` -- drop-in for the reference VQ layer (/root/reference/vqvae.py:28-78).

Same constructor, attributes, registered buffers (names, order, shapes, dtype), return tuple and
autograd behaviour as the reference class, so `VQVAE.quantize_t / quantize_b` (vqvae.py:185,190),
`VQVAE_Deep` (vqvae_deep.py:252,257), train_vqvae.py and extract_code.py can use it unchanged and
reference checkpoints load strictly.  All arithmetic runs in hand-written sm_100a CUDA behind the C
ABI of include/vqb200.h; PyTorch supplies device memory, streams and torch.distributed only.
There is no CPU or PyTorch fallback: non-CUDA tensors raise.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
from torch import nn

from . import _native
from . import distributed as dist_fn


def _collapse(sizes, strides):
    """Collapse dims into one (size, stride) if they are nested-contiguous; None otherwise."""
    dims = [(s, st) for s, st in zip(sizes, strides) if s != 1]
    if not dims:
        return 1, 0
    for (s0, st0), (s1, st1) in zip(dims[:-1], dims[1:]):
        if st0 != st1 * s1:
            return None
    n = 1
    for s, _ in dims:
        n *= s
    return n, dims[-1][1]


_LAYOUT_CACHE = {}
_TC_CACHE = {}
_ENV_PIN_FILTER = bool(os.environ.get("VQB200_TC_SPLIT"))      # read once: the filter precision is pinned by the environment


def _tc_supported(lib, ptr, lay, dim, n_embed):
    """vqb200_tc_supported, memoised on (layout, shape, pointer alignment)."""
    key = (lay, dim, n_embed, ptr & 15)
    hit = _TC_CACHE.get(key)
    if hit is None:
        n, rpi, img, row, col = lay
        hit = bool(lib.vqb200_tc_supported(ptr, n, dim, n_embed, rpi, img, row, col))
        if len(_TC_CACHE) < 4096:
            _TC_CACHE[key] = hit
    return hit


def _raw_stream(dev):
    return torch._C._cuda_getCurrentRawStream(dev.index)


def row_layout(x: torch.Tensor):
    """Map a [..., D] tensor to the C ABI's row layout (n_rows, rows_per_image, image_stride,
    row_stride, col_stride) without copying, or None when the strides fit neither accepted form."""
    key = (tuple(x.shape), x.stride())
    hit = _LAYOUT_CACHE.get(key, 0)
    if hit != 0:
        return hit
    lay = _row_layout(x.shape, x.stride())
    if len(_LAYOUT_CACHE) < 4096:
        _LAYOUT_CACHE[key] = lay
    return lay


def _row_layout(shape, stride):
    D = shape[-1]
    lead_sizes, lead_strides = list(shape[:-1]), list(stride[:-1])
    col = stride[-1] if D > 1 else 1
    n = 1
    for s in lead_sizes:
        n *= s
    if n == 0:
        return 0, 1, 0, D, 1
    if col <= 0:
        return None
    if n == 1:                                # a single row: any row stride describes it
        return 1, 1, 0, (D if col == 1 else 1), col
    for split in range(len(lead_sizes) + 1):
        outer = _collapse(lead_sizes[:split], lead_strides[:split])
        inner = _collapse(lead_sizes[split:], lead_strides[split:])
        if outer is None or inner is None:
            continue
        (n_img, img_stride), (rpi, row_stride) = outer, inner
        if rpi == 1:
            rpi, row_stride, n_img, img_stride = n_img, img_stride, 1, 0
        if row_stride <= 0 or (n_img > 1 and img_stride <= 0):
            continue
        if col != 1 and row_stride != 1:
            continue
        return n, rpi, (img_stride if n_img > 1 else 0), row_stride, col
    return None


def _non_overlapping_and_dense(x: torch.Tensor) -> bool:
    """True when x's elements tile one gap-free block of memory in SOME dimension order (then torch's element-wise ops,
    and vqvae.py:73 with them, return a tensor with exactly x's strides)."""
    if x.is_contiguous():                    # (the common case, one C++ call)
        return True
    dims = sorted((st, n) for n, st in zip(x.shape, x.stride()) if n != 1)
    expect = 1
    for st, n in dims:
        if st != expect:
            return False
        expect *= n
    return True


class _QuantizeFunction(torch.autograd.Function):
    """forward = CUDA kernels; backward = straight-through + commitment gradient (vqvae.py:72-73)."""

    @staticmethod
    def forward(ctx, x, module):
        quantize, diff, ind, image, lay = module._run_forward(x, keep_image=x.requires_grad)
        ctx.save_for_backward(x, ind)
        ctx.image, ctx.lay, ctx.dims = image, lay, (module.dim, module.n_embed)
        ctx.mark_non_differentiable(ind)
        return quantize, diff, ind

    @staticmethod
    def backward(ctx, grad_quantize, grad_diff, _grad_ind):
        x, ind = ctx.saved_tensors
        lib = _native.load()
        n, rpi, img, row, col = ctx.lay
        dim, n_embed = ctx.dims
        gq = None
        if grad_quantize is not None:
            gq = grad_quantize
            if gq.stride() != x.stride() or gq.dtype != torch.float32:
                gq = torch.empty_strided(x.shape, x.stride(), dtype=torch.float32, device=x.device)
                gq.copy_(grad_quantize)
        gd = None
        if grad_diff is not None:
            gd = grad_diff.to(torch.float32).reshape(1).contiguous()
        gx = torch.empty_strided(x.shape, x.stride(), dtype=torch.float32, device=x.device)
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        with torch.cuda.device(x.device):
            _native.check(lib.vqb200_quantize_backward(
                _native.ptr(x), _native.ptr(ind), _native.ptr(ctx.image), _native.ptr(gq), _native.ptr(gd),
                _native.ptr(gx), n, dim, n_embed, rpi, img, row, col, stream), "vqb200_quantize_backward")
        return gx, None


class Quantize(nn.Module):
    """Vector-quantisation layer with EMA codebook (reference vqvae.py:28-78).

    Extra keyword (not in the reference, defaults keep reference behaviour):
      engine: "auto" | "simt" | "tcgen05" | "tcgen05_bf16" -- which assignment kernel the C ABI uses; "auto"
              picks the tcgen05 engine when the shape is covered and adapts its filter precision (see _pick_engine).
    """

    def __init__(self, dim, n_embed, decay=0.99, eps=1e-5, engine="auto"):
        super().__init__()
        self.dim = dim
        self.n_embed = n_embed
        self.decay = decay
        self.eps = eps
        if engine not in _native.ENGINES:
            raise ValueError(f"engine must be one of {sorted(_native.ENGINES)}")
        self.engine = engine
        self.check_ids = False

        embed = torch.randn(dim, n_embed)                                 # vqvae.py:37
        self.register_buffer("embed", embed)                              # vqvae.py:38
        self.register_buffer("cluster_size", torch.zeros(n_embed))        # vqvae.py:39
        self.register_buffer("embed_avg", embed.clone())                  # vqvae.py:40
        # private device workspaces (never in the state_dict)
        self._ws = {}
        # adaptive filter precision of the tcgen05 engine (engine="auto"): speed only, results are exact either way
        self._filter = {"mode": "bf16", "cooldown": 0, "pending": None}

    def __getstate__(self):
        # device workspaces / CUDA events are neither pickled nor deep-copied; they are rebuilt on first use
        state = self.__dict__.copy()
        state["_ws"] = {}
        state["_filter"] = {"mode": "bf16", "cooldown": 0, "pending": None}
        return state

    # ------------------------------------------------------------------ workspaces
    def _workspace(self, device, n_rows):
        lib = _native.load()
        ws = self._ws.get(device)
        if ws is None:
            ws = {"image": torch.empty(lib.vqb200_codebook_bytes(self.dim, self.n_embed), dtype=torch.uint8, device=device),
                  "stats": torch.zeros(lib.vqb200_stats_bytes(self.dim, self.n_embed) // 4, dtype=torch.float32, device=device),
                  "scratch": None, "rows": -1}
            self._ws = {device: ws}           # one device at a time (module.to() moves it)
        if ws["rows"] < n_rows:
            ws["scratch"] = torch.empty(lib.vqb200_forward_scratch_bytes(n_rows, self.dim, self.n_embed),
                                        dtype=torch.uint8, device=device)
            ws["rows"] = n_rows
        return ws

    # ------------------------------------------------------------------ peer memory for the fused all-reduce
    def _peer_workspace(self, ws, dev):
        """Symmetric (peer-mapped) receive buffers of all ranks, set up collectively on the first multi-rank training forward:
        per step parity, one receive slot per rank for the packed statistics (every rank's fold + EMA kernel stores its
        statistics as {value, step} pairs into slot `rank` of every rank's buffer and polls its own, local, slots).  Returns None (-> NCCL all-reduce + vqb200_ema_update) when peer memory is unavailable, the world is
        larger than 8 ranks, the shape is outside the fused EMA kernel, or VQB200_NO_P2P is set."""
        if "peer" in ws:
            return ws["peer"]
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("Quantize: the first multi-rank training forward sets up peer memory collectively and cannot "
                               "be captured -- run one eager forward on every rank before capturing the CUDA graph")
        ws["peer"] = None
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        # fused into the fold + EMA kernel at dim 64 / n_embed 256, 512; a separate exchange kernel for every other shape
        # whose packed statistics fit the slots (<= 1 Mi words: up to D = 256, K = 4096 ... D = 64, K = 16384)
        ok = (world <= 8 and self.n_embed * (self.dim + 1) <= (1 << 20) and not os.environ.get("VQB200_NO_P2P"))
        peer = None
        if ok:
            try:
                import torch.distributed._symmetric_memory as symm
                n = self.n_embed * (self.dim + 1)    # packed statistics words [K*D sums | K counts]
                n_al = (n + 63) // 64 * 64
                slot = 2 * n_al                      # floats per receive slot: every word travels as an 8-byte {value, step} pair
                slots = world * slot
                total = 2 * slots + 64               # [slots parity 0 | slots parity 1 | time-out (2) / trace words ... step counter (+32)]
                buf = symm.empty(total, dtype=torch.float32, device=dev)
                hdl = symm.rendezvous(buf, dist.group.WORLD)
                buf.zero_()
                torch.cuda.synchronize(dev)
                hdl.barrier()
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                mk = lambda vals: (C.c_void_p * (2 * world))(*vals)
                peer = {"buf": buf, "hdl": hdl, "rank": rank, "world": world, "calls": 0, "checked": 0, "words": n,
                        "fused": self.dim == 64 and self.n_embed in (256, 512),
                        # [parity][rank r]: where MY statistics go on rank r (peer-mapped) ...
                        "push_dst": mk([p + 4 * (par * slots + rank * slot) for par in (0, 1) for p in ptrs]),
                        # ... and what the kernel polls and sums: my LOCAL slots, one per rank
                        "recv": mk([ptrs[rank] + 4 * (par * slots + r * slot) for par in (0, 1) for r in range(world)]),
                        "err_ptr": ptrs[rank] + 4 * (2 * slots),            # 16 words: time-out record (+ trace words)
                        "step_ptr": ptrs[rank] + 4 * (2 * slots + 32),      # the kernel's own exchange counter
                        "err": buf[2 * slots: 2 * slots + 2].view(torch.int32)}
            except Exception as exc:  # no peer access / unsupported build: keep the NCCL path
                peer = None
                self._peer_error = repr(exc)
        # every rank must take the same path: agree collectively
        flag = torch.tensor([1 if peer is not None else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            peer = None
        ws["peer"] = peer
        return peer

    def _check_peer_timeout(self, peer):
        """The EMA kernel gives up after 2 s without a peer's statistics and records (step, rank) in its flag array; read
        the word back every 64 steps without a sync (pinned copy + event) and fail loudly."""
        pend = peer.get("err_pending")
        if pend is not None and pend[0].query():
            peer["err_pending"] = None
            if int(pend[1][0]) != 0:
                raise RuntimeError(f"Quantize: rank {int(pend[1][1])} did not publish its codebook statistics for step "
                                   f"{int(pend[1][0])} within 2 s (fused peer-memory exchange); the replicas are out of sync")
        if pend is None and peer["calls"] - peer["checked"] >= 64 and not torch.cuda.is_current_stream_capturing():
            peer["checked"] = peer["calls"]
            host = peer.setdefault("err_host", torch.zeros(2, dtype=torch.int32).pin_memory())
            host.copy_(peer["err"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            peer["err_pending"] = (ev, host)

    # ------------------------------------------------------------------ filter precision policy
    # The plain-bf16 tensor-core filter needs 1/3 of the MMAs of the split-bf16 one but certifies fewer rows when
    # the two nearest codes are almost equidistant (e.g. N(0,1) inputs against a random codebook); uncertified rows
    # cost an exact fp32 re-score.  The number of re-scored rows of a call is read back asynchronously (no sync) and
    # decides the precision of later calls.  Either way the returned indices are the exact arg-min.
    FLAG_SWITCH_FRACTION = 1e-2               # (the fix-up of <= 1 % of the rows costs less than the split filter's extra MMAs)
    SPLIT_COOLDOWN_CALLS = 64
    FLAG_SAMPLE_EVERY = 8                     # bf16-mode calls between two read-backs of the counter

    def _pick_engine(self, x, lay):
        if self.engine != "auto" or _ENV_PIN_FILTER:
            return _native.ENGINES[self.engine]
        if lay[0] == 0 or not _tc_supported(_native.load(), x.data_ptr(), lay, self.dim, self.n_embed):
            return _native.ENGINE_AUTO
        f = self._filter
        pend = f["pending"]
        if pend is not None and torch.cuda.is_current_stream_capturing():   # CUDA-graph capture: no event queries, keep the filter as it is
            return _native.ENGINE_TCGEN05_BF16 if f["mode"] == "bf16" else _native.ENGINE_TCGEN05
        if pend is not None and pend[0].query():
            _, host_count, rows, mode_used = pend
            f["pending"] = None
            if int(host_count[10]) != 0:      # scratch header byte 56: the statistics kernel met an index outside [0, n_embed)
                raise RuntimeError("Quantize: internal error -- the code-statistics kernel read an out-of-range index")
            if mode_used == "bf16" and int(host_count[0]) > self.FLAG_SWITCH_FRACTION * rows:
                f["mode"], f["cooldown"] = "split", self.SPLIT_COOLDOWN_CALLS
        if f["mode"] == "split":
            f["cooldown"] -= 1
            if f["cooldown"] <= 0:
                f["mode"] = "bf16"                # probe the cheap filter again
        return _native.ENGINE_TCGEN05_BF16 if f["mode"] == "bf16" else _native.ENGINE_TCGEN05

    def _note_flagged(self, ws, n, eng, dev):
        f = self._filter
        if eng != _native.ENGINE_TCGEN05_BF16 or self.engine != "auto" or f["pending"] is not None or n == 0:
            return
        f["calls"] = f.get("calls", 0) + 1
        if f["calls"] % self.FLAG_SAMPLE_EVERY != 1:
            return
        if torch.cuda.is_current_stream_capturing():      # no pinned read-back / event inside a captured forward
            return
        host = ws.get("flag_host")
        if host is None:
            host = ws["flag_host"] = torch.zeros(11, dtype=torch.int32).pin_memory()
        # vqb200.h: scratch header -- int32 at byte 16 = rows sent to the exact re-score, int32 at byte 56 = internal-error flag
        host.copy_(ws["scratch"][16:60].view(torch.int32), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        f["pending"] = (ev, host, n, "bf16" if eng == _native.ENGINE_TCGEN05_BF16 else "split")

    def _check_input(self, x):
        if not isinstance(x, torch.Tensor):
            raise TypeError("Quantize expects a torch.Tensor")
        if x.dim() < 1 or x.shape[-1] != self.dim:
            raise RuntimeError(f"Quantize: last dimension of input {tuple(x.shape)} must equal dim={self.dim}")
        if x.dtype != torch.float32:
            raise RuntimeError(f"Quantize: expected float32 input (the codebook is float32), got {x.dtype}")
        if not x.is_cuda:
            raise RuntimeError("Quantize (B200-native): input must be a CUDA tensor; there is no CPU fallback")
        if self.embed.device != x.device or self.embed.dtype != torch.float32:
            raise RuntimeError("Quantize: module buffers must be float32 on the input's device")

    # ------------------------------------------------------------------ forward
    def _run_forward(self, x, keep_image=False, want_quantize=True, lay=None):
        lib = _native.load()
        if lay is None:
            lay = row_layout(x)
        if lay is None:
            raise RuntimeError("Quantize: unsupported input strides (internal: forward() copies such inputs)")
        n, rpi, img, row, col = lay
        dev = x.device
        ws = self._workspace(dev, n)
        image = ws["image"]
        if keep_image:                        # backward gathers from the codebook this forward used
            image = torch.empty_like(ws["image"])
        if torch.cuda.current_device() != dev.index:
            torch.cuda.set_device(dev)        # kernels launch on the input's device
        stream = _raw_stream(dev)
        quantize = torch.empty_strided(x.shape, x.stride(), dtype=torch.float32, device=dev) if want_quantize else None
        ind = torch.empty(x.shape[:-1], dtype=torch.int64, device=dev)
        diff = torch.empty((), dtype=torch.float32, device=dev)
        stats = ws["stats"] if self.training else None
        world = dist_fn.get_world_size() if self.training else 1
        peer = None
        if world > 1:
            peer = self._peer_workspace(ws, dev)
            if peer is not None:              # this step's statistics go straight into every rank's receive slot
                self._check_peer_timeout(peer)
                peer["calls"] += 1            # (the step tag of the exchange lives on the device: nothing here can fall out of step)
        bufs = self._buffers                  # (nn.Module.__getattr__ costs ~0.3 us per access)
        embed, cluster_size, embed_avg = bufs["embed"], bufs["cluster_size"], bufs["embed_avg"]
        x_run, q_run, lay_run, x_dense = x, quantize, lay, None
        strided = n > 0 and (col != 1 or (n > 1 and row != self.dim))
        if strided and self.engine != "simt":
            if _tc_supported(lib, x.data_ptr(), lay, self.dim, self.n_embed):
                # NCHW-physical rows the tensor-core kernel consumes in place; in training mode its converters also write
                # the dense copy of x the code-statistics kernel gathers from
                if self.training:
                    x_dense = torch.empty((n, self.dim), dtype=torch.float32, device=dev)
            elif _tc_supported(lib, 0, (n, n, 0, self.dim, 1), self.dim, self.n_embed):
                # other strided layouts on a covered shape: re-pack to dense rows (coalesced CUDA transpose), run the
                # tensor-core engine, re-pack `quantize` back to the input's strides (vqvae.py:73)
                x_run = torch.empty((n, self.dim), dtype=torch.float32, device=dev)
                _native.check(lib.vqb200_repack_rows(x.data_ptr(), x_run.data_ptr(), n, self.dim, rpi, img, row, col, 1, stream),
                              "vqb200_repack_rows")
                q_run = torch.empty_like(x_run) if want_quantize else None
                lay_run = (n, n, 0, self.dim, 1)
        eng = self._pick_engine(x_run, lay_run)
        n, rpi, img, row, col = lay_run
        fused_ema = self.training and world == 1
        # the codebook image is re-derived from `embed` on every call: external writes to the buffer
        # (load_state_dict, .data.copy_, DDP buffer broadcast) can never leave it stale
        if peer is not None and peer["fused"]:    # forward + statistics pushed to every rank (vqvae.py:43-56,58-59,72-73)
            _native.check(lib.vqb200_quantize_step_peers(
                x_run.data_ptr(), n, self.dim, self.n_embed, rpi, img, row, col, embed.data_ptr(),
                cluster_size.data_ptr(), embed_avg.data_ptr(), image.data_ptr(),
                q_run.data_ptr() if q_run is not None else None, ind.data_ptr(), diff.data_ptr(), ws["scratch"].data_ptr(),
                x_dense.data_ptr() if x_dense is not None else None, eng, float(self.decay), float(1 - self.decay), float(self.eps),
                peer["push_dst"], peer["recv"], peer["err_ptr"], peer["step_ptr"], peer["rank"], peer["world"], stream),
                "vqb200_quantize_step_peers")
        else:
            _native.check(lib.vqb200_quantize_step(
                x_run.data_ptr(), n, self.dim, self.n_embed, rpi, img, row, col, embed.data_ptr(),
                cluster_size.data_ptr(), embed_avg.data_ptr(), image.data_ptr(),
                q_run.data_ptr() if q_run is not None else None, ind.data_ptr(), diff.data_ptr(),
                stats.data_ptr() if stats is not None else None, ws["scratch"].data_ptr(),
                x_dense.data_ptr() if x_dense is not None else None, eng, 1 if fused_ema else 0, float(self.decay), float(1 - self.decay), float(self.eps), stream),
                "vqb200_quantize_step")                                                       # vqvae.py:43-73
        if q_run is not quantize:             # dense result -> the input's strides
            _native.check(lib.vqb200_repack_rows(q_run.data_ptr(), quantize.data_ptr(), lay[0], self.dim, lay[1], lay[2],
                                                 lay[3], lay[4], 0, stream), "vqb200_repack_rows")
        self._note_flagged(ws, n, eng, dev)
        if self.training and not fused_ema and not (peer is not None and peer["fused"]):
            if peer is not None:              # vqvae.py:58-59 as an in-place exchange over peer memory (any shape)
                _native.check(lib.vqb200_stats_exchange_peers(
                    stats.data_ptr(), peer["words"], peer["push_dst"], peer["recv"], peer["err_ptr"], peer["step_ptr"],
                    peer["rank"], peer["world"], stream), "vqb200_stats_exchange_peers")
            else:                             # no peer memory: one packed NCCL all-reduce
                dist_fn.all_reduce(stats[: self.n_embed * (self.dim + 1)])
            _native.check(lib.vqb200_ema_update(
                stats.data_ptr(), cluster_size.data_ptr(), embed_avg.data_ptr(),
                embed.data_ptr(), self.dim, self.n_embed, float(self.decay), float(1 - self.decay),
                float(self.eps), None, stream), "vqb200_ema_update")                          # vqvae.py:61-70
        return quantize, diff, ind, image, lay

    def forward(self, input):
        self._check_input(input)
        out_like = None
        if not _non_overlapping_and_dense(input):
            # strided slices, expanded views: vqvae.py:73 (`input + (...)`) returns a DENSE tensor in the input's dimension
            # order -- exactly what clone(preserve_format) allocates; the kernels then work on / return that layout
            input = input.clone(memory_format=torch.preserve_format)
        lay = row_layout(input)
        if lay is None:                       # dense, but in a dimension order the kernels do not address: one explicit
            out_like = input                  # copy, like reshape() in vqvae.py:43, and the result goes back to that order
            input = input.contiguous()
        if input.requires_grad and torch.is_grad_enabled():
            quantize, diff, embed_ind = _QuantizeFunction.apply(input, self)
        else:
            quantize, diff, embed_ind, _, _ = self._run_forward(input, lay=lay)
        if out_like is not None:
            quantize = torch.empty_like(out_like).copy_(quantize)
        return quantize, diff, embed_ind

    @torch.no_grad()
    def assign(self, input):
        """Index extraction only (extract_code.py:23 uses nothing but the ids): no quantize / diff output,
        never touches the EMA buffers."""
        self._check_input(input)
        if row_layout(input) is None:
            input = input.contiguous()
        was = self.training
        self.training = False
        try:
            _, _, ind, _, _ = self._run_forward(input, want_quantize=False)
        finally:
            self.training = was
        return ind

    def embed_code(self, embed_id):
        """vqvae.py:77-78: F.embedding(embed_id, embed.T) -> [..., dim]."""
        if not embed_id.is_cuda:
            raise RuntimeError("Quantize.embed_code (B200-native): embed_id must be a CUDA tensor")
        if embed_id.dtype != torch.int64:
            if embed_id.dtype not in (torch.int32,):
                raise RuntimeError("Quantize.embed_code: expected an integer index tensor (int64/int32)")
            embed_id = embed_id.to(torch.int64)
        lib = _native.load()
        dev = embed_id.device
        ids = embed_id.contiguous()
        n = ids.numel()
        ws = self._workspace(dev, 0)
        out = torch.empty(tuple(embed_id.shape) + (self.dim,), dtype=torch.float32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _native.check(lib.vqb200_codebook_prepare(_native.ptr(self.embed), self.dim, self.n_embed,
                                                      _native.ptr(ws["image"]), stream), "vqb200_codebook_prepare")
            _native.check(lib.vqb200_embed_code(_native.ptr(ids), n, _native.ptr(ws["image"]), self.dim,
                                                self.n_embed, _native.ptr(out), _native.ptr(status), stream),
                          "vqb200_embed_code")
        if self.check_ids and int(status.item()) != 0:   # opt-in (costs a host sync): F.embedding-style range check
            raise IndexError("Quantize.embed_code: index out of range")
        return out
