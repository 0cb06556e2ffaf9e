#!/usr/bin/env python
"""Benchmark of the B200-native `Quantize` hot path (BASELINE.json metric: quantized vectors/s, fwd+EMA).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one training-mode forward (assignment + gather + commitment loss + statistics + EMA update)
of the bottom quantizer of the 256-px VQ-VAE on one batch of synthetic latents: cfg-2 of BASELINE.json,
x = [128, 64, 64, 64] fp32 (N = 524 288 vectors, D = 64, K = 512) PER GPU (weak scaling; the codebook
statistics are all-reduced over NCCL as one packed buffer).  Rank 0 prints ONE JSON line.

  value     : vectors/s with inputs resident in HBM (device timed, CUDA events, max over ranks)
  e2e       : the same step through the host-buffer C ABI (vqb200_host_quantize): pinned host x in,
              quantize + indices + diff back to pinned host memory, copies inside the timed region
  roofline  : HBM roofline of the fused forward call (algorithmic bytes N*(8D+8))
  cpu_baseline / --impl reference : the CPU oracle port of the reference algorithm on the host cores
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
CPU_SAMPLE_ROWS = 8 * 64 * 64                  # cfg-1 bottom latent: what the CPU path is timed on


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


def time_cpu_port(steps, warmup):
    """The oracle (numpy restatement of vqvae.py:42-75, OpenBLAS on all host cores), train mode, fwd+EMA."""
    from oracle.quantize_oracle import QuantizeOracle
    embed, x = synth_inputs_numpy(CPU_SAMPLE_ROWS, 1234)
    o = QuantizeOracle(D, K, embed=embed)
    x = x.reshape(8, 64, 64, D)
    for _ in range(max(warmup, 1)):
        o.forward(x)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        o.forward(x)
        ts.append(time.perf_counter() - t0)
    med = float(np.median(ts))
    return CPU_SAMPLE_ROWS / med, med


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
    """--impl reference: the reference's CPU implementation of the path (oracle port; the PyTorch reference
    itself cannot travel to the GPU box), all host threads, bounded sample per step."""
    if rank != 0:
        return
    vps, med = time_cpu_port(args.steps, args.warmup)
    cores = os.cpu_count() or 1
    sample = f"{CPU_SAMPLE_ROWS} rows (cfg-1 bottom latent [8,64,64,64]) per step, numpy/OpenBLAS fp32, train mode"
    line = {"impl": "reference", "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg-2 bottom quantizer D=64 K=512 fwd+EMA (bounded CPU sample)",
                       "rows_per_step": CPU_SAMPLE_ROWS, "dim": D, "n_embed": K},
            "cpu_baseline": {"value": vps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": vps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tcgen05", "tcgen05_bf16"])
    ap.add_argument("--dist", default="clustered", choices=["clustered", "randn"])
    ap.add_argument("--layout", default="dense", choices=["dense", "nchw"],
                    help="dense: contiguous [B,H,W,D] rows (cfg-2 primary); nchw: the permute(0,2,3,1) view VQVAE.encode passes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _native.load()
    steps, warmup = args.steps, max(args.warmup, 3)

    # ---- synthetic workload: 3 rotating input batches (134 MB each > 126 MB L2, and never the same twice in a row)
    torch.manual_seed(0)
    q = vq.Quantize(D, K, engine=args.engine).to(dev).train()
    embed0 = q.embed.clone()
    xs = []
    for i in range(3):
        g = torch.Generator(device=dev).manual_seed(1234 + 1000 * i + rank)
        if args.dist == "clustered":
            pick = torch.randint(0, K, (N_ROWS,), device=dev, generator=g)
            x = embed0.t()[pick] + 0.1 * torch.randn(N_ROWS, D, device=dev, generator=g)
        else:
            x = torch.randn(N_ROWS, D, device=dev, generator=g)
        x = x.reshape(B, H, W, D).contiguous()
        if args.layout == "nchw":
            x = x.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
        xs.append(x)

    def step(i):
        return q(xs[i % 3])

    def reset():
        # EMA steady state of a trained codebook: cluster_size = expected rows per code, embed_avg = embed * cluster_size,
        # so training steps keep the codes where the (clustered) data are instead of replaying the reference's
        # start-up transient (cluster_size starts at 0 -> never-hit codes blow up ~1e5x after the first update)
        q.embed.data.copy_(embed0)
        if args.dist == "clustered":
            q.cluster_size.data.fill_(float(world * N_ROWS) / K)
            q.embed_avg.data.copy_(embed0 * (float(world * N_ROWS) / K))
        else:
            q.embed_avg.data.copy_(embed0); q.cluster_size.data.zero_()

    try:
        gpu_uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local_rank, gpu_uuid)
    if rank == 0:
        sampler.start()                       # before the warm-up: the poller is running long before the timed region
    reset()
    for i in range(warmup):
        step(i)
    reset()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if world > 1:
        # ranks leave the NCCL barrier hundreds of microseconds apart on the host; the step itself keeps them in lock-step
        # (peer-memory flags), so the earliest starter would time the others' head start.  All ranks of the node share
        # the host clock: agree on a start instant a few milliseconds ahead and spin until it.
        t_go = torch.tensor([time.time() + 0.004], dtype=torch.float64, device=dev)
        dist.broadcast(t_go, 0)
        t_go = float(t_go.item())
        while time.time() < t_go:
            pass
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.vqb200_launch_count()
    ev0.record()
    t_host0 = time.perf_counter()
    for i in range(steps):
        step(i)
    host_issue_ms = (time.perf_counter() - t_host0) * 1e3 / steps      # host time to ISSUE a step (no sync inside)
    ev1.record()
    gpu_launches = int(lib.vqb200_launch_count() - launches0)
    torch.cuda.synchronize()
    t_wall1 = time.time()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    elapsed_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / steps
    value = world * N_ROWS / (ms_per_step * 1e-3)

    # ---- roofline of the fused forward call (the dominant launch), timed alone with events on its stream
    peak, peak_src = measured_peaks()
    ws = q._workspace(dev, N_ROWS)
    quant = torch.empty(B, H, W, D, device=dev); ind = torch.empty(B, H, W, dtype=torch.int64, device=dev)
    diff = torch.empty((), device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    eng = _native.ENGINES[args.engine]
    if args.engine == "auto":                 # time the kernel variant the module's precision policy settled on
        eng = _native.ENGINE_TCGEN05_BF16 if q._filter["mode"] == "bf16" else _native.ENGINE_TCGEN05
    reset()
    _native.check(lib.vqb200_codebook_prepare(_native.ptr(q.embed), D, K, _native.ptr(ws["image"]), stream), "prepare")

    xd = xs if args.layout == "dense" else [x.contiguous() for x in xs]      # the kernel-level timing uses dense rows

    def fwd_only(i, with_stats):
        _native.check(lib.vqb200_quantize_forward(_native.ptr(xd[i % 3]), N_ROWS, D, K, N_ROWS, 0, D, 1,
                                                  _native.ptr(ws["image"]), _native.ptr(quant), _native.ptr(ind),
                                                  _native.ptr(diff), _native.ptr(ws["stats"]) if with_stats else None,
                                                  _native.ptr(ws["scratch"]), eng, stream), "forward")

    def time_fwd(with_stats):
        for i in range(3):
            fwd_only(i, with_stats)
        torch.cuda.synchronize()
        ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ka.record()
        for i in range(steps):
            fwd_only(i, with_stats)
        kb.record()
        torch.cuda.synchronize()
        return ka.elapsed_time(kb) / steps

    # dominant kernel = tc::k_vq_tc (assignment + gather + straight-through value + loss), timed ALONE with CUDA events
    # on its stream (vqb200_debug_tc_kernel launches nothing else); the SIMT engine has no single dominant launch, so
    # there the whole forward call is timed
    fwd_ms = time_fwd(False)
    fwd_stats_ms = time_fwd(True)
    kernel_ms, kernel_name = fwd_ms, "vqb200_quantize_forward (SIMT engine: k_assign_exact + k_gather_stats)"
    if eng in (_native.ENGINE_TCGEN05, _native.ENGINE_TCGEN05_BF16):
        ws["scratch"][:256].zero_()

        def kern_only(i):
            _native.check(lib.vqb200_debug_tc_kernel(_native.ptr(xd[i % 3]), N_ROWS, D, K, _native.ptr(ws["image"]),
                                                     _native.ptr(quant), _native.ptr(ind), _native.ptr(ws["scratch"]),
                                                     eng, stream), "tc_kernel")
        for i in range(3):
            kern_only(i)
        torch.cuda.synchronize()
        ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ka.record()
        for i in range(steps):
            kern_only(i)
        kb.record()
        torch.cuda.synchronize()
        kernel_ms = ka.elapsed_time(kb) / steps
        kernel_name = ("tc::k_vq_tc<%s> (tcgen05 distance filter + certified arg-min + gather + straight-through value + loss)"
                       % ("plain bf16" if eng == _native.ENGINE_TCGEN05_BF16 else "split bf16, CTA pair"))
    algo_bytes = N_ROWS * (8 * D + 8)
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_dram_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("k_vq_tc_dram_bytes_per_launch")
        except Exception:
            traffic = None
    tflops = 2.0 * N_ROWS * D * K / (kernel_ms * 1e-3) / 1e12
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel_name, "launch_ms": kernel_ms, "algorithmic_bytes": algo_bytes,
                "peak_source": peak_src, "tensor_flops": 2.0 * N_ROWS * D * K, "tensor_tflops_achieved": tflops,
                "tensor_frac_of_measured_bf16_peak": tflops / 1659.1,
                "forward_call_ms": fwd_ms, "statistics_kernels_ms": max(fwd_stats_ms - fwd_ms, 0.0),
                "statistics_algorithmic_bytes": N_ROWS * (4 * D + 8)}

    # ---- e2e through the host-buffer C ABI (pinned host in / out, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        reset()
        torch.cuda.synchronize()
        ctx = C.c_void_p()
        _native.check(lib.vqb200_host_ctx_create(N_ROWS, D, K, C.byref(ctx)), "host_ctx_create")
        hx = [x.reshape(N_ROWS, D).cpu().pin_memory() for x in xs[:2]]
        hq = torch.empty(N_ROWS, D).pin_memory()
        hi = torch.empty(N_ROWS, dtype=torch.int64).pin_memory()
        hd = torch.empty(1).pin_memory()

        def host_step(i):
            _native.check(lib.vqb200_host_quantize(ctx, C.c_void_p(hx[i % 2].data_ptr()), N_ROWS, _native.ptr(q.embed),
                                                   _native.ptr(q.cluster_size), _native.ptr(q.embed_avg), 0.99,
                                                   float(1 - 0.99), 1e-5, 1, C.c_void_p(hq.data_ptr()),
                                                   C.c_void_p(hi.data_ptr()), C.c_void_p(hd.data_ptr()), eng), "host_quantize")
        e_steps = max(5, min(steps, 20))
        for i in range(3):
            host_step(i)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(e_steps):
            host_step(i)
        dt = time.perf_counter() - t0
        lib.vqb200_host_ctx_destroy(ctx)
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * N_ROWS * e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": N_ROWS * D * 4,
               "d2h_bytes_per_step": N_ROWS * D * 4 + N_ROWS * 8 + 4, "ms_per_step": dt / e_steps * 1e3,
               "steps": e_steps, "api": "vqb200_host_quantize (C ABI, pinned host buffers, stats not all-reduced)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        vps, med = time_cpu_port(7, 2)
        cpu = {"value": vps, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"{CPU_SAMPLE_ROWS} rows (cfg-1 bottom latent [8,64,64,64]), median of 7, numpy/OpenBLAS fp32 oracle"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "cfg-2: bottom quantizer, x=[128,64,64,64] fp32 per GPU, D=64, K=512, train fwd+EMA",
                           "rows_per_gpu": N_ROWS, "dim": D, "n_embed": K, "distribution": args.dist, "engine": args.engine, "layout": args.layout,
                           "l2_policy": "3 rotating 134 MB input batches (each larger than the 126 MB L2)",
                           "codebook_state": "EMA steady state (cluster_size = N/K, embed_avg = embed*N/K)" if args.dist == "clustered" else "reference init",
                           "parallelism": f"dp{world}",
                           "collective": ("none" if world == 1 else ("all-reduce fused into the EMA kernel over peer memory (NVLink P2P loads)"
                                                                      if q._ws.get(dev, {}).get("peer") is not None else "NCCL all-reduce of the packed statistics"))},
                "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "host_issue_ms_per_step": host_issue_ms,
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
