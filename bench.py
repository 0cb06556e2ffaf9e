#!/usr/bin/env python
"""Benchmark of the B200-native `Quantize` hot path (BASELINE.json metric: quantized vectors/s, fwd+EMA).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one training-mode forward (assignment + gather + commitment loss + statistics + EMA update)
of the bottom quantizer of the 256-px VQ-VAE on one batch of synthetic latents: cfg-2 of BASELINE.json,
x = [128, 64, 64, 64] fp32 (N = 524 288 vectors, D = 64, K = 512) PER GPU (weak scaling; the codebook
statistics are all-reduced over NCCL as one packed buffer).  Rank 0 prints ONE JSON line.

  value     : vectors/s with inputs resident in HBM (device timed, CUDA events, max over ranks; the K-step region is
              measured in several back-to-back windows, each bracketed as the contract says, and the MEDIAN window is
              reported -- one host hiccup on any rank must not decide a 2-ms measurement; all windows are in `timing`)
  e2e       : the same step through the host-buffer C ABI (vqb200_host_quantize) at N = 1 -- pinned host x in,
              quantize + indices + diff back to pinned host memory, copies inside the timed region; at N > 1 through the
              module itself (pinned host x -> device -> forward with the cross-rank exchange -> pinned host results)
  roofline  : HBM roofline of the dominant kernel (algorithmic bytes N*(8D+8)), timed alone with CUDA events
  cpu_baseline / --impl reference : the reference's own `Quantize` (oracle/_ref/vqvae.py, staged unmodified by
              tools/fetch_ref.py) on the host cores, thread count set explicitly and reported
  reference_gpu : the same reference module on the B200 (TF32 off) on the same inputs, N = 1 only
  parity_multi  : N > 1 only -- after the timed loop: replicas bit-identical, fused peer-memory exchange == NCCL exchange,
              multi-rank result == the single-rank result on the concatenated batch; non-zero exit on a mismatch
  nchw / cfg3   : secondary measurements (the layout VQVAE.encode passes; top+bottom step at global batch 256)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, H, W, D, K = 128, 64, 64, 64, 512          # cfg-2 (per GPU)
N_ROWS = B * H * W
METRIC = "quantized vectors/sec (fwd+EMA)"
UNIT = "vectors/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def synth_inputs_numpy(rows, seed):
    """clustered distribution (ii) of SURVEY 8d: embed[:, randint(K)] + 0.1 N(0,1); embed = randn under seed 0."""
    import torch
    torch.manual_seed(0)
    embed = torch.randn(D, K)
    g = torch.Generator().manual_seed(seed)
    pick = torch.randint(0, K, (rows,), generator=g)
    x = embed.t()[pick] + 0.1 * torch.randn(rows, D, generator=g)
    return embed.numpy().copy(), x.numpy().copy()


def load_reference_quantize():
    """The reference's own Quantize class (oracle/_ref/vqvae.py, unmodified, digests checked); None if not staged."""
    try:
        from oracle import reference_module
        return reference_module.load("vqvae").Quantize
    except Exception:
        return None


def time_cpu_reference(steps, warmup, rows, threads=None):
    """Reference `Quantize(64, 512).train()` forward (+EMA) on the host cores, clustered inputs, `rows` vectors per step.
    Falls back to the numpy port of the same algorithm (oracle/quantize_oracle.py) when the reference is not staged.
    Returns (vectors/s, median seconds per step, kind, threads)."""
    import torch
    threads = threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)             # torchrun exports OMP_NUM_THREADS=1: set explicitly, and report it
    embed, x = synth_inputs_numpy(rows, 1234)
    ref_cls = load_reference_quantize()
    if ref_cls is not None:
        torch.manual_seed(0)
        q = ref_cls(D, K).train()
        with torch.no_grad():
            q.embed.copy_(torch.from_numpy(embed)); q.embed_avg.copy_(torch.from_numpy(embed) * (rows / K))
            q.cluster_size.fill_(rows / K)
        xt = torch.from_numpy(x).reshape(-1, 64, 64, D)
        fn, kind = (lambda: q(xt)), "reference"
    else:
        from oracle.quantize_oracle import QuantizeOracle
        o = QuantizeOracle(D, K, embed=embed)
        xr = x.reshape(-1, 64, 64, D)
        fn, kind = (lambda: o.forward(xr)), "port"
    with torch.no_grad():
        for _ in range(max(warmup, 1)):
            fn()
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
    med = float(np.median(ts))
    return rows / med, med, kind, threads


_NVML_POLL = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
key = sys.argv[1]
try:
    h = nv.nvmlDeviceGetHandleByUUID(key if key.startswith("GPU-") else "GPU-" + key) if len(key) > 8 else nv.nvmlDeviceGetHandleByIndex(int(key))
except Exception:
    h = nv.nvmlDeviceGetHandleByIndex(0)
smax = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
bits = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
        ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
print("ready", smax, flush=True)
out = sys.stdout
while True:
    t = time.time()
    sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
    out.write("%.6f %d %s\n" % (t, sm, ",".join(n for n, b in bits if r & b) or "-"))
    out.flush()
    time.sleep(0.0005)
"""


class ClockSampler:
    """SM clock and throttle reasons of one GPU, polled from a SEPARATE process (no GIL contention with the issuing
    loop) through NVML about once per millisecond, every sample time-stamped; stop(t0, t1) keeps the samples taken
    inside the timed region [t0, t1] (host clock, both ends after a device synchronize).  Falls back to
    `nvidia-smi -lms` (the recipe's clocks line) when NVML is not importable."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, uuid=None):
        self.index, self.uuid, self.rows, self.proc, self.mode, self.smax = index, uuid, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _NVML_POLL, str(self.uuid or self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            head = self.proc.stdout.readline().split()
            if len(head) != 2 or head[0] != "ready":
                raise RuntimeError("nvml poller did not start")
            self.smax, self.mode = float(head[1]), "nvml"
        except Exception:
            try:
                if self.proc:
                    self.proc.kill()
                self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                              "--format=csv,noheader,nounits", "-lms", "100"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.mode = "nvidia-smi"
            except Exception:
                self.proc = None
                return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line)

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock sampler unavailable"]}
        time.sleep(0.02 if self.mode == "nvml" else 0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = list(self.rows)
        window = "timed region" if (self.mode == "nvml" and t0 is not None) else "sampler lifetime"
        if self.mode == "nvml" and t0 is not None:
            inside = [ln for ln in rows if ln.split() and t0 <= float(ln.split()[0]) <= t1]
            if len(inside) < 3:                  # very short timed region: take the loaded neighbourhood (warm-up .. roofline runs)
                t0, t1, window = t0 - 0.05, t1 + 0.05, "timed region +-50 ms"
        sm, smax, reasons, n_all = [], [], set(), 0
        for line in rows:
            try:
                if self.mode == "nvml":
                    ts, mhz, why = line.split()
                    n_all += 1
                    if t0 is not None and not (t0 <= float(ts) <= t1):
                        continue
                    sm.append(float(mhz)); smax.append(self.smax)
                    reasons.update(w for w in why.split(",") if w != "-")
                else:
                    r = [c.strip() for c in line.split(",")]
                    sm.append(float(r[1])); smax.append(float(r[2])); n_all += 1
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": self.mode,
                "window": window}


def run_reference_arm(args, rank):
    """--impl reference: the reference's CPU implementation of the path -- its own `Quantize` module from oracle/_ref on all
    host threads.  Workload = cfg-2 itself (524 288 rows per step) when K steps of it fit in about two minutes, else the
    largest power-of-two fraction of it that does (stated in `sample`)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rows = N_ROWS
    _, probe, kind, threads = time_cpu_reference(1, 1, 32768)
    budget_s = 120.0
    while rows > 32768 and probe * (rows / 32768) * (args.steps + max(args.warmup, 1)) > budget_s:
        rows //= 2
    vps, med, kind, threads = time_cpu_reference(args.steps, args.warmup, rows, cores)
    what = "the reference's own torch Quantize (oracle/_ref/vqvae.py, unmodified)" if kind == "reference" else "numpy port of vqvae.py:42-75"
    sample = (f"{rows} rows per step ([{rows // 4096},64,64,64] of cfg-2's [128,64,64,64]), {what}, fp32, train mode, "
              f"{threads} torch threads, clustered inputs")
    line = {"impl": "reference", "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg-2 bottom quantizer D=64 K=512 fwd+EMA on the host CPU" + ("" if rows == N_ROWS else " (bounded sample)"),
                       "rows_per_step": rows, "dim": D, "n_embed": K, "same_config": rows == N_ROWS},
            "cpu_baseline": {"value": vps, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": vps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def timed_windows(step, steps, windows, world, dev, dist):
    """`windows` back-to-back measurements of EXACTLY `steps` steps, each bracketed by barrier + synchronize on both sides
    and timed with CUDA events on the launching stream; per window the MAX over ranks.  Returns (list of ms, wall t0, wall t1)."""
    import torch
    out = []
    t_first = t_last = None
    for w in range(windows):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
            # ranks leave the NCCL barrier hundreds of microseconds apart on the host while the step keeps them in lock-step
            # (peer-memory flags): agree on a start instant on the node's shared host clock and spin until it
            t_go = torch.tensor([time.time() + 0.003], dtype=torch.float64, device=dev)
            dist.broadcast(t_go, 0)
            t_go = float(t_go.item())
            while time.time() < t_go:
                pass
        t0 = time.time()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            step(i)
        ev1.record()
        torch.cuda.synchronize()
        t1 = time.time()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        out.append(ms)
        t_first = t0 if t_first is None else t_first
        t_last = t1
    return out, t_first, t_last


def scaled_max_err(a, b, scale):
    import torch
    d = (a.double() - b.double()).abs()
    return float(torch.where(d == 0, torch.zeros_like(d), d / scale).max())


def parity_multi(vq, dev, rank, world, dist):
    """N > 1: the exchange of vqvae.py:58-59.  Two training steps on rank-specific inputs through (a) the default multi-rank
    path (all-reduce fused into the EMA kernel over peer memory when available) and (b) the NCCL all-reduce path; then
    every rank's buffers are gathered: replicas must be bit-identical, (a) must equal (b), and both must equal the
    SINGLE-rank module run on the concatenated batch (the single-rank path is what the GPU test-suite pins to the reference)."""
    import torch
    rows = 16384
    res = {"rows_per_rank": rows, "steps": 2}
    torch.manual_seed(0)
    ref = vq.Quantize(D, K).to(dev).train()                         # single-rank twin, fed the concatenated batch
    mods = {}
    for name in ("default", "nccl"):
        if name == "nccl":
            os.environ["VQB200_NO_P2P"] = "1"
        m = vq.Quantize(D, K).to(dev).train()
        m.load_state_dict(ref.state_dict())
        mods[name] = m
    os.environ.pop("VQB200_NO_P2P", None)
    import vq_vae_2_pytorch_b200.distributed as dist_fn
    worst = {"default_vs_nccl": 0.0, "multi_vs_single_rank": 0.0}
    identical = True
    for step in range(2):
        g = torch.Generator(device=dev).manual_seed(4242 + 97 * step + rank)
        live = torch.nonzero(ref.embed.pow(2).sum(0) < 100.0 * ref.embed.pow(2).sum(0).min()).reshape(-1)
        pick = live[torch.randint(0, live.numel(), (rows,), device=dev, generator=g)]
        x = (ref.embed.t()[pick] + (0.1 if step else 1.0) * torch.randn(rows, D, device=dev, generator=g)).contiguous()
        allx = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(allx, x)
        before_avg = ref.embed_avg.clone()
        outs = {name: m(x) for name, m in mods.items()}
        world_saved = dist_fn.get_world_size
        dist_fn.get_world_size = lambda group=None: 1              # the twin runs the single-rank path on all rows
        try:
            xcat = torch.cat(allx)
            _, _, ind_cat = ref(xcat)
        finally:
            dist_fn.get_world_size = world_saved
        torch.cuda.synchronize()
        sabs = torch.zeros(K, D, dtype=torch.float64, device=dev).index_add_(0, ind_cat, xcat.abs().double())
        scale = (0.99 * before_avg.double().abs() + 0.01 * sabs.t()).clamp_min(1e-30)
        for name, m in mods.items():
            assert torch.equal(outs[name][2], ind_cat[rank * rows:(rank + 1) * rows])
            gathered = [torch.empty_like(m.embed_avg) for _ in range(world)]
            dist.all_gather(gathered, m.embed_avg)
            identical &= all(torch.equal(gathered[0], t) for t in gathered[1:])
            gathered = [torch.empty_like(m.embed) for _ in range(world)]
            dist.all_gather(gathered, m.embed)
            identical &= all(torch.equal(gathered[0], t) for t in gathered[1:])
            worst["multi_vs_single_rank"] = max(worst["multi_vs_single_rank"], scaled_max_err(m.embed_avg, ref.embed_avg, scale),
                                                scaled_max_err(m.cluster_size, ref.cluster_size, ref.cluster_size.double().abs().clamp_min(1e-30)))
        worst["default_vs_nccl"] = max(worst["default_vs_nccl"], scaled_max_err(mods["default"].embed_avg, mods["nccl"].embed_avg, scale))
        # continue from ONE common state: the single-rank twin's summation order is not reproducible between processes
        # (shared-memory adds of buckets that straddle two warps), so rank 0's twin is broadcast before it is loaded
        for buf in (ref.embed, ref.cluster_size, ref.embed_avg):
            dist.broadcast(buf, 0)
        for m in mods.values():
            m.load_state_dict(ref.state_dict())
    res.update(worst)
    res["replicas_bit_identical"] = bool(identical)
    res["default_path"] = ("fused peer-memory exchange" if mods["default"]._ws.get(dev, {}).get("peer") is not None else "NCCL all-reduce")
    flag = torch.tensor([1 if (identical and worst["default_vs_nccl"] <= 1e-5 and worst["multi_vs_single_rank"] <= 1e-5) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["ok"] = bool(int(flag.item()))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tcgen05", "tcgen05_bf16", "tcgen05_tf32"])
    ap.add_argument("--dist", default="clustered", choices=["clustered", "randn"])
    ap.add_argument("--layout", default="dense", choices=["dense", "nchw"],
                    help="dense: contiguous [B,H,W,D] rows (cfg-2 primary); nchw: the permute(0,2,3,1) view VQVAE.encode passes")
    ap.add_argument("--windows", type=int, default=5, help="back-to-back measurements of the K-step region; the median is reported")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip reference_gpu / nchw / cfg3 / parity_multi")
    ap.add_argument("--cfg5", action="store_true", help="also run the cfg-5 codebook sweep (K = 512 .. 8192, D = 64 .. 256, N = 524 288 per GPU, train fwd+EMA)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    import vq_vae_2_pytorch_b200 as vq
    from vq_vae_2_pytorch_b200 import _native

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # several ranks share one host: keep this rank's threads -- and with them its pinned buffers (first touch) -- on the CPUs
        # / NUMA node next to its GPU, so that the e2e copies do not cross the socket interconnect
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByUUID(str(torch.cuda.get_device_properties(dev).uuid).encode()
                                            if hasattr(torch.cuda.get_device_properties(dev), "uuid") else b""))
        except Exception:
            try:
                import pynvml
                pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            except Exception:
                pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _native.load()
    steps, warmup = args.steps, max(args.warmup, 3)

    # ---- synthetic workload: 3 rotating input batches (134 MB each > 126 MB L2, and never the same twice in a row)
    torch.manual_seed(0)
    q = vq.Quantize(D, K, engine=args.engine).to(dev).train()
    embed0 = q.embed.clone()

    def make_batches(rows, shape, layout, seed0):
        out = []
        for i in range(3):
            g = torch.Generator(device=dev).manual_seed(seed0 + 1000 * i + rank)
            if args.dist == "clustered":
                pick = torch.randint(0, K, (rows,), device=dev, generator=g)
                x = embed0.t()[pick] + 0.1 * torch.randn(rows, D, device=dev, generator=g)
            else:
                x = torch.randn(rows, D, device=dev, generator=g)
            x = x.reshape(shape).contiguous()
            if layout == "nchw":
                x = x.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
            out.append(x)
        return out

    xs = make_batches(N_ROWS, (B, H, W, D), args.layout, 1234)

    def reset(mod, rows_total):
        # EMA steady state of a trained codebook: cluster_size = expected rows per code, embed_avg = embed * cluster_size,
        # so training steps keep the codes where the (clustered) data are instead of replaying the reference's
        # start-up transient (cluster_size starts at 0 -> never-hit codes blow up ~1e5x after the first update)
        mod.embed.data.copy_(embed0)
        if args.dist == "clustered":
            mod.cluster_size.data.fill_(float(rows_total) / K)
            mod.embed_avg.data.copy_(embed0 * (float(rows_total) / K))
        else:
            mod.embed_avg.data.copy_(embed0); mod.cluster_size.data.zero_()

    def step(i):
        return q(xs[i % 3])

    try:
        gpu_uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local_rank, gpu_uuid)
    if rank == 0:
        sampler.start()                       # before the warm-up: the poller is running long before the timed region
    reset(q, world * N_ROWS)
    for i in range(warmup):
        step(i)
    reset(q, world * N_ROWS)
    torch.cuda.synchronize()
    launches0 = lib.vqb200_launch_count()
    t_host0 = time.perf_counter()
    for i in range(steps):                    # one untimed pass: host time to ISSUE a step (no sync inside) and launches per step
        step(i)
    host_issue_ms = (time.perf_counter() - t_host0) * 1e3 / steps
    gpu_launches = int(lib.vqb200_launch_count() - launches0)
    windows_ms, t_wall0, t_wall1 = timed_windows(step, steps, max(1, args.windows), world, dev, dist)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    elapsed_ms = float(np.median(windows_ms))
    ms_per_step = elapsed_ms / steps
    value = world * N_ROWS / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel, timed alone with events on its stream
    peak, peak_src = measured_peaks()
    ws = q._workspace(dev, N_ROWS)
    quant = torch.empty(B, H, W, D, device=dev); ind = torch.empty(B, H, W, dtype=torch.int64, device=dev)
    diff = torch.empty((), device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    eng = _native.ENGINES[args.engine]
    if args.engine == "auto":                 # time the kernel variant the module's precision policy settled on
        eng = _native.ENGINE_TCGEN05_BF16 if q._filter["mode"] == "bf16" else _native.ENGINE_TCGEN05
    reset(q, world * N_ROWS)
    _native.check(lib.vqb200_codebook_prepare(_native.ptr(q.embed), D, K, _native.ptr(ws["image"]), stream), "prepare")

    xd = xs if args.layout == "dense" else [x.contiguous() for x in xs]      # the kernel-level timing uses dense rows

    def fwd_only(i, with_stats):
        _native.check(lib.vqb200_quantize_forward(_native.ptr(xd[i % 3]), N_ROWS, D, K, N_ROWS, 0, D, 1,
                                                  _native.ptr(ws["image"]), _native.ptr(quant), _native.ptr(ind),
                                                  _native.ptr(diff), _native.ptr(ws["stats"]) if with_stats else None,
                                                  _native.ptr(ws["scratch"]), eng, stream), "forward")

    def time_fn(fn, n):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ka.record()
        for i in range(n):
            fn(i)
        kb.record()
        torch.cuda.synchronize()
        return ka.elapsed_time(kb) / n

    # dominant kernel = tc::k_vq_tc (assignment + gather + straight-through value + loss), timed ALONE with CUDA events
    # on its stream (vqb200_debug_tc_kernel launches nothing else); the SIMT engine has no single dominant launch, so
    # there the whole forward call is timed
    fwd_ms = time_fn(lambda i: fwd_only(i, False), steps)
    fwd_stats_ms = time_fn(lambda i: fwd_only(i, True), steps)
    kernel_ms, kernel_name = fwd_ms, "vqb200_quantize_forward (SIMT engine: k_assign_exact + k_gather_stats)"
    tc_engines = (_native.ENGINE_TCGEN05, _native.ENGINE_TCGEN05_BF16, _native.ENGINE_TCGEN05_TF32)
    if eng in tc_engines:
        ws["scratch"][:256].zero_()

        def kern_only(i):
            _native.check(lib.vqb200_debug_tc_kernel(_native.ptr(xd[i % 3]), N_ROWS, D, K, _native.ptr(ws["image"]),
                                                     _native.ptr(quant), _native.ptr(ind), _native.ptr(ws["scratch"]),
                                                     eng, stream), "tc_kernel")
        kernel_ms = time_fn(kern_only, steps)
        kernel_name = ("tc::k_vq_tc<%s> (tcgen05 distance filter + certified arg-min + gather + straight-through value + loss)"
                       % {_native.ENGINE_TCGEN05_BF16: "plain bf16", _native.ENGINE_TCGEN05: "split bf16, CTA pair",
                          _native.ENGINE_TCGEN05_TF32: "tf32 from the fp32 stage"}[eng])
    algo_bytes = N_ROWS * (8 * D + 8)
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    for name in ("r02_dram_traffic.json", "r01_dram_traffic.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("k_vq_tc_dram_bytes_per_launch")
                traffic_src = f"static: profiles/{name} (one `ncu --set full` capture of this kernel; not measured in this run)"
                break
            except Exception:
                traffic = None
    tflops = 2.0 * N_ROWS * D * K / (kernel_ms * 1e-3) / 1e12
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name, "launch_ms": kernel_ms,
                "algorithmic_bytes": algo_bytes, "peak_source": peak_src, "tensor_flops": 2.0 * N_ROWS * D * K,
                "tensor_tflops_achieved": tflops, "tensor_frac_of_measured_bf16_peak": tflops / 1659.1,
                "forward_call_ms": fwd_ms, "statistics_kernels_ms": max(fwd_stats_ms - fwd_ms, 0.0),
                "statistics_algorithmic_bytes": N_ROWS * (4 * D + 8),
                "step_level": {"algorithmic_bytes": algo_bytes, "ms_per_step": ms_per_step,
                               "achieved": algo_bytes / (ms_per_step * 1e-3) / 1e9, "frac": algo_bytes / (ms_per_step * 1e-3) / 1e9 / peak}}

    # ---- e2e
    e2e = None
    if not args.no_e2e:
        reset(q, world * N_ROWS)
        torch.cuda.synchronize()
        hx = [x.reshape(N_ROWS, D).cpu().pin_memory() for x in xs[:2]]
        hq = torch.empty(N_ROWS, D).pin_memory()
        hi = torch.empty(N_ROWS, dtype=torch.int64).pin_memory()
        hd = torch.empty(1).pin_memory()
        e_steps = max(5, min(steps, 20))
        # through the host-buffer C ABI (pinned host in / out, copies inside the timed region, chunks overlapped)
        ctx = C.c_void_p()
        _native.check(lib.vqb200_host_ctx_create(N_ROWS, D, K, C.byref(ctx)), "host_ctx_create")
        if world == 1:
            def host_step(i):
                _native.check(lib.vqb200_host_quantize(ctx, C.c_void_p(hx[i % 2].data_ptr()), N_ROWS, _native.ptr(q.embed),
                                                       _native.ptr(q.cluster_size), _native.ptr(q.embed_avg), 0.99,
                                                       float(1 - 0.99), 1e-5, 1, C.c_void_p(hq.data_ptr()),
                                                       C.c_void_p(hi.data_ptr()), C.c_void_p(hd.data_ptr()), eng), "host_quantize")
            api = "vqb200_host_quantize (C ABI, pinned host buffers, row chunks pipelined over three streams)"
        else:
            # the data-parallel form of the same call: statistics left on the device, reduced where the reference calls
            # dist_fn.all_reduce (vqvae.py:58-59), then the EMA of vqvae.py:61-70 -- all inside the timed region
            stats = torch.zeros(lib.vqb200_stats_bytes(D, K) // 4, device=dev)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _native.check(lib.vqb200_host_ctx_set_stream(ctx, st), "host_ctx_set_stream")

            def host_step(i):
                _native.check(lib.vqb200_host_quantize_stats(ctx, C.c_void_p(hx[i % 2].data_ptr()), N_ROWS, _native.ptr(q.embed),
                                                             _native.ptr(stats), C.c_void_p(hq.data_ptr()),
                                                             C.c_void_p(hi.data_ptr()), C.c_void_p(hd.data_ptr()), eng), "host_quantize_stats")
                dist.all_reduce(stats[: K * (D + 1)])
                _native.check(lib.vqb200_ema_update(_native.ptr(stats), _native.ptr(q.cluster_size), _native.ptr(q.embed_avg),
                                                    _native.ptr(q.embed), D, K, 0.99, float(1 - 0.99), 1e-5, None, st), "ema_update")
                torch.cuda.synchronize()
            api = ("vqb200_host_quantize_stats (C ABI, pinned host buffers, row chunks pipelined over three streams) + one NCCL "
                   "all-reduce of the packed statistics + vqb200_ema_update, per rank")
        for i in range(3):
            host_step(i)
        if world > 1:
            dist.barrier()
        # three back-to-back windows of e_steps calls each (wall clock: the call returns when the results are in host memory;
        # max over ranks per window); the median window is reported, all three are listed
        e_windows = []
        for w in range(3):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for i in range(e_steps):
                host_step(i)
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            e_windows.append(dt)
        lib.vqb200_host_ctx_destroy(ctx)
        dt = sorted(e_windows)[1]
        e2e = {"value": world * N_ROWS * e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": N_ROWS * D * 4,
               "d2h_bytes_per_step": N_ROWS * D * 4 + N_ROWS * 8 + 4, "ms_per_step": dt / e_steps * 1e3,
               "steps": e_steps, "window_ms": [w * 1e3 for w in e_windows], "api": api}

    extras = {}
    if not args.no_extras:
        x_steps = max(5, min(steps, 20))
        # ---- the layout VQVAE.encode really passes (vqvae.py:227,235): same workload, NCHW-physical rows
        if args.layout == "dense":
            xs_n = [x.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1) for x in xs]
            reset(q, world * N_ROWS)
            for i in range(3):
                q(xs_n[i])
            reset(q, world * N_ROWS)
            w_ms, _, _ = timed_windows(lambda i: q(xs_n[i % 3]), x_steps, 3, world, dev, dist)
            ms = float(np.median(w_ms)) / x_steps
            extras["nchw"] = {"ms_per_step": ms, "value": world * N_ROWS / (ms * 1e-3), "unit": UNIT,
                              "layout": "permute(0,2,3,1) view of an NCHW tensor, consumed and written in place"}
            del xs_n
        # ---- cfg-3: one optimiser step's worth = top [b,32,32,64] then bottom [b,64,64,64], global batch 256 split as 256 / n_gpu
        bsz = 256 // world
        qt = vq.Quantize(D, K, engine=args.engine).to(dev).train()
        qb = vq.Quantize(D, K, engine=args.engine).to(dev).train()
        xt = make_batches(bsz * 1024, (bsz, 32, 32, D), "nchw", 777)
        xb = make_batches(bsz * 4096, (bsz, 64, 64, D), "nchw", 888)
        reset(qt, 256 * 1024); reset(qb, 256 * 4096)

        def step3(i):
            qt(xt[i % 3]); qb(xb[i % 3])
        for i in range(3):
            step3(i)
        reset(qt, 256 * 1024); reset(qb, 256 * 4096)
        w_ms, _, _ = timed_windows(step3, x_steps, 3, world, dev, dist)
        ms = float(np.median(w_ms)) / x_steps
        extras["cfg3"] = {"ms_per_step": ms, "value": 256 * 5120 / (ms * 1e-3), "unit": UNIT, "scaling": "strong",
                          "workload": f"top [b,32,32,64] + bottom [b,64,64,64] NCHW-physical, train fwd+EMA, global batch 256 = {world} x {bsz}"}
        del xt, xb, qt, qb
        # ---- the reference's own Quantize on this GPU (TF32 off), same inputs -- N = 1 only (it is a single-process number)
        if world == 1 and rank == 0:
            ref_cls = load_reference_quantize()
            if ref_cls is not None:
                torch.backends.cuda.matmul.allow_tf32 = False
                r = ref_cls(D, K).to(dev).train()
                reset(r, N_ROWS)
                xr = make_batches(N_ROWS, (B, H, W, D), "dense", 1234)
                with torch.no_grad():
                    for i in range(2):
                        r(xr[i])
                    reset(r, N_ROWS)
                    ms = time_fn(lambda i: r(xr[i % 3]), 5)
                extras["reference_gpu"] = {"ms_per_step": ms, "value": N_ROWS / (ms * 1e-3), "unit": UNIT,
                                           "what": "the reference's own torch Quantize (oracle/_ref/vqvae.py, unmodified) on this B200, fp32 (TF32 off), same cfg-2 inputs, train fwd+EMA"}
                del r, xr
                torch.cuda.empty_cache()
        if world > 1:
            extras["parity_multi"] = parity_multi(vq, dev, rank, world, dist)
    if args.cfg5:
        # cfg-5: scaled codebook sweep, batch 128 (N = 524 288 rows) per GPU, training forward + EMA (statistics all-reduced
        # across ranks over peer memory: fused into the EMA kernel at D = 64 / K = 512, the separate in-place exchange kernel elsewhere)
        sweep = []
        for d5 in (64, 128, 256):
            for k5 in (512, 1024, 2048, 4096, 8192):
                torch.manual_seed(0)
                m5 = vq.Quantize(d5, k5).to(dev).train()
                e5 = m5.embed.clone()
                g5 = torch.Generator(device=dev).manual_seed(55 + rank)
                pick = torch.randint(0, k5, (N_ROWS,), device=dev, generator=g5)
                x5 = (e5.t()[pick] + 0.1 * torch.randn(N_ROWS, d5, device=dev, generator=g5)).contiguous()
                m5.cluster_size.data.fill_(float(world * N_ROWS) / k5); m5.embed_avg.data.copy_(e5 * (float(world * N_ROWS) / k5))
                for _ in range(3):
                    m5(x5)
                w_ms, _, _ = timed_windows(lambda i: m5(x5), 5, 3, world, dev, dist)
                ms = float(np.median(w_ms)) / 5
                sweep.append({"dim": d5, "n_embed": k5, "ms_per_step": ms, "value": world * N_ROWS / (ms * 1e-3),
                              "tensor_tflops_algorithmic": 2.0 * world * N_ROWS * d5 * k5 / (ms * 1e-3) / 1e12})
                del m5, x5
                torch.cuda.empty_cache()
        extras["cfg5"] = {"unit": UNIT, "rows_per_gpu": N_ROWS, "mode": "train fwd+EMA, clustered rows", "points": sweep}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rows = 131072                            # [32,64,64,64]: a quarter of cfg-2 per step, ~10 s of CPU work in all
        vps, med, kind, threads = time_cpu_reference(5, 1, rows)
        cpu = {"value": vps, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"{rows} rows per step ([32,64,64,64], a quarter of cfg-2), median of 5, "
                         + ("the reference's own torch Quantize (oracle/_ref, unmodified)" if kind == "reference" else "numpy port (reference not staged)")
                         + f", fp32, train mode, {threads} torch threads"}

    if rank == 0:
        peer = q._ws.get(dev, {}).get("peer") if world > 1 else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "cfg-2: bottom quantizer, x=[128,64,64,64] fp32 per GPU, D=64, K=512, train fwd+EMA",
                           "rows_per_gpu": N_ROWS, "dim": D, "n_embed": K, "distribution": args.dist, "engine": args.engine, "layout": args.layout,
                           "l2_policy": "3 rotating 134 MB input batches (each larger than the 126 MB L2)",
                           "codebook_state": "EMA steady state (cluster_size = N/K, embed_avg = embed*N/K)" if args.dist == "clustered" else "reference init",
                           "parallelism": f"dp{world}",
                           "collective": ("none" if world == 1 else ("all-reduce fused into the EMA kernel over peer memory (NVLink P2P loads)"
                                                                      if peer is not None else "NCCL all-reduce of the packed statistics"))},
                "timing": {"windows": len(windows_ms), "window_ms": windows_ms, "reported": "median window / steps; each window = exactly `steps` steps, barrier + synchronize on both sides, CUDA events, max over ranks"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "host_issue_ms_per_step": host_issue_ms,
                "roofline": roofline, "cpu_baseline": cpu}
        line.update(extras)
        print(json.dumps(line), flush=True)
    ok = extras.get("parity_multi", {}).get("ok", True)
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("bench.py: multi-rank parity check failed (see parity_multi)")


if __name__ == "__main__":
    main()
