#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native LFM sampling path.

    python bench.py --gpus N --steps K --warmup W            # ours (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on the host CPU

Workload (BASELINE.json configs[1]): DiT-L/2, 32x32x4 latents, Euler-50 (torchdiffeq-style uniform grid,
step_size 0.02 => 50 NFE), batch 64 per GPU, synthetic non-degenerate weights (lfm_b200.synthetic
seed 1 - the reference's own init is all-zero at the output) and Gaussian latents.  One "step" = one full
sampling pass over one batch of 64 latents per GPU.  Metric: images/s (whole job, weak scaling).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "DiT-L/2"
BATCH = 64
NFE = 50
STEP_SIZE = 1.0 / NFE
WEIGHT_SEED = 1


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(burst=float(d["bf16_tflops"]), sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thr = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()  # the exact child we started
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                power.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [c for c in sm if smax and c > 0.3 * smax] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm)}


def build_model(device):
    import lfm_b200
    from lfm_b200.synthetic import synthetic_state_dict
    with torch.device("meta"):
        net = lfm_b200.DiT_models[MODEL](img_resolution=32, in_channels=4, label_dropout=0.0, num_classes=1)
    sd = synthetic_state_dict(net, WEIGHT_SEED)
    net = net.to_empty(device="cpu")
    net.load_state_dict(sd, strict=True)
    net = net.to(device)
    net.max_batch_hint = BATCH
    return net


def time_dominant_kernel(device, peaks, iters=20):
    """Roofline of the dominant kernel: the CTA-pair tcgen05 GEMM (gemm2_bf16_tcgen05), on its largest instance
    in the network (mlp.fc1 + bias + GELU: M = 64*256, N = 4096, K = 1024), timed with CUDA events on the
    launching stream.  Operands + output = 176 MB > the 126 MB L2."""
    from lfm_b200 import _lib
    lib = _lib.load()
    M, N, K = BATCH * 256, 4096, 1024
    a = torch.randn(M, K, device=device).bfloat16()
    w = (torch.randn(N, K, device=device) * 0.03).bfloat16()
    bias = torch.randn(N, device=device)
    out = torch.empty(M, N, device=device, dtype=torch.bfloat16)
    s = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)

    def run():
        rc = lib.lfm_dbg_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), None, 0, 256, M, N, K, 1, 512, s)
        assert rc == 0, _lib.last_error()
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / iters
    achieved = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("gemm_fc1_dram_bytes_per_launch")
    return {"bound": "tensor", "kernel": "gemm2_bf16_tcgen05<EPI_BIAS_GELU_BF16> (cta_group::2, 256x256 tile) M=16384 N=4096 K=1024",
            "achieved": round(achieved, 1), "peak": peaks["burst"], "unit": "TFLOP/s", "frac": round(achieved / peaks["burst"], 4),
            "traffic": traffic, "us_per_launch": round(ms * 1e3, 2), "peak_source": peaks["source"] + ", burst (kernel timed alone)"}


def usable_threads():
    """Host threads the CPU legs should use.  os.cpu_count() reports the machine (128 on the GPU box) while the
    container may be entitled to far fewer cores; oversubscribing MKL there is ~50x slower.  Calibrate: time a
    1024^3 fp32 matmul at a few thread counts and keep the fastest."""
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    a = torch.randn(1024, 1024)
    best, best_t = 1, float("inf")
    cand = sorted({c for c in (4, 8, 16, 32, 64, 128, avail) if c <= avail})
    for c in cand:
        torch.set_num_threads(c)
        a @ a
        t0 = time.perf_counter()
        for _ in range(3):
            a @ a
        dt = time.perf_counter() - t0
        if dt < best_t * 0.9:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_baseline(threads, seconds_hint=20):
    """The oracle (CPU port of the reference algorithm, fp32) on the host cores, bounded sample."""
    from oracle import dit as odit
    from oracle import solvers as osol
    torch.set_num_threads(threads)
    cfg = odit.make_config(MODEL, num_classes=1, label_dropout=0.0)
    sd = odit.synthetic_state_dict(cfg, WEIGHT_SEED)
    B = 4   # the CPU is most efficient at small batches (measured: batch 16 halves its NFE*img/s)
    x = torch.randn(B, 4, 32, 32, generator=torch.Generator().manual_seed(0))
    f = lambda t, xx: odit.dit_forward(sd, cfg, t, xx)  # noqa: E731
    osol.tdq_euler(f, x, 1.0)  # warm-up: 1 NFE
    t0 = time.time()
    _, nfe = osol.tdq_euler(f, x, 0.04)  # bounded sample of the workload: 25 of the 50 Euler steps, batch 4 (~10-20 s)
    dt = time.time() - t0
    return B * nfe / dt / NFE, dt, f"DiT-L/2 fp32 oracle, batch {B}, {nfe} Euler NFE in {dt:.2f}s, scaled to Euler-{NFE} images/s"


def workload_config(world):
    """The `config` object of both arms (the driver compares them)."""
    return {"workload": f"{MODEL} 32x32x4 latents, Euler-{NFE} (step_size {STEP_SIZE}), batch {BATCH}/GPU, "
                        f"synthetic non-degenerate init (seed {WEIGHT_SEED}), {'dp' + str(world)}",
            "l2": "inputs larger than L2: 0.9 GB of weights + 0.7 GB of activations stream per NFE (L2 = 126 MB)",
            "global_batch": BATCH * world, "nfe_per_image": NFE}


def run_reference(args):
    """--impl reference: the reference algorithm (oracle port; the reference is pure Python and cannot travel to
    the GPU box) on the host CPU with all the threads it can use; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = usable_threads()
    from oracle import dit as odit
    from oracle import solvers as osol
    cfg = odit.make_config(MODEL, num_classes=1, label_dropout=0.0)
    sd = odit.synthetic_state_dict(cfg, WEIGHT_SEED)
    B, nfe_s = 4, 10
    x = torch.randn(B, 4, 32, 32, generator=torch.Generator().manual_seed(0))
    f = lambda t, xx: odit.dit_forward(sd, cfg, t, xx)  # noqa: E731
    for _ in range(max(1, args.warmup)):
        osol.tdq_euler(f, x, 1.0)
    t0 = time.time()
    for _ in range(args.steps):
        osol.tdq_euler(f, x, 1.0 / nfe_s)
    dt = (time.time() - t0) / args.steps
    value = B * nfe_s / dt / NFE
    sample = f"each step = DiT-L/2 fp32, batch {B}, {nfe_s} Euler NFE on {threads} threads; images/s scaled to Euler-{NFE}"
    line = {"impl": "reference", "metric": "images/sec DiT-L/2 32x32 latents Euler-50", "value": value, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(int(os.environ.get("WORLD_SIZE", "1"))),
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    import lfm_b200
    from lfm_b200 import dist as ldist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); the lfm_b200 hot path has no CPU fallback")
    rank, world, local = ldist.init_from_env("nccl")
    if world != args.gpus and not (world == 1 and args.gpus == 1):
        if rank == 0:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    peaks = load_peaks()
    net = build_model(device)
    flops_nfe = 161_386_856_448  # algorithmic FLOPs per sample per NFE, DiT-L/2
    sargs = types.SimpleNamespace(method="euler", step_size=STEP_SIZE, perturb=False, cfg_scale=1.0, compute_nfe=False)
    chw = 4 * 32 * 32

    # host-side buffers for the end-to-end leg (pinned)
    g = torch.Generator().manual_seed(ldist.rank_seed(42, rank))
    z_host = torch.randn(BATCH, 4, 32, 32, generator=g).pin_memory()
    out_host = torch.empty(BATCH, 4, 32, 32).pin_memory()
    z_dev = z_host.to(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def step_device():
        xf = lfm_b200.sample_from_model(net, z_dev, {}, sargs)[-1]
        return ldist.all_gather_batch(xf) if world > 1 else xf

    def step_e2e():
        z = z_host.to(device, non_blocking=True)
        xf = lfm_b200.sample_from_model(net, z, {}, sargs)[-1]
        allx = ldist.all_gather_batch(xf) if world > 1 else xf
        out_host.copy_(allx[rank::world] if world > 1 else allx, non_blocking=True)
        return allx

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = net.launch_count()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), net.launch_count() - l0

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, launches = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    if rank == 0:
        imgs = BATCH * world * args.steps
        value = imgs / (ms_total * 1e-3)
        e2e_value = imgs / (ms_e2e * 1e-3)
        tflops = value * NFE * flops_nfe / 1e12 / world  # per GPU
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = usable_threads()
            v, dt, sample = cpu_baseline(threads)
            cpu = {"value": round(v, 5), "unit": "images/s", "cores": threads, "kind": "port", "sample": sample}
        elif world == 1:
            time.sleep(3.0)
        # the dominant kernel is timed ALONE against the burst peak: the CPU leg above (or the pause) lets the board
        # leave the power-capped state of the long run first, as in the measurement of the burst peak itself
        roof = time_dominant_kernel(device, peaks) if world == 1 else None
        line = {
            "metric": "images/sec DiT-L/2 32x32 latents Euler-50", "value": round(value, 3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world),
            "nfe_img_per_s": round(value * NFE, 1),
            "e2e": {"value": round(e2e_value, 3), "unit": "images/s", "h2d_bytes_per_step": BATCH * chw * 4,
                    "d2h_bytes_per_step": BATCH * chw * 4, "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline_forward": {"bound": "tensor", "achieved": round(tflops, 1), "unit": "TFLOP/s per GPU",
                                 "peak_sustained": peaks["sustained"], "frac_sustained": round(tflops / peaks["sustained"], 4),
                                 "peak_burst": peaks["burst"], "frac_burst": round(tflops / peaks["burst"], 4),
                                 "flops_per_sample_per_nfe": flops_nfe, "peak_source": peaks["source"]},
        }
        if roof is not None:
            line["roofline"] = roof
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
