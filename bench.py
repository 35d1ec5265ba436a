#!/usr/bin/env python
"""Benchmark of the 3D-DDPM sampling hot path on B200 (contract: see the task statement / DESIGN.md).

A "step" is one DDPM reverse step of one 96^3 low-dose-conditioned patch with the paper-default
UNet (BASELINE.json configs[1]): one UNet evaluation + the posterior / noise update.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

* `value`  : UNet evals/s, whole job, inputs resident in HBM, the device-resident loop
             (ddpm3d_sample_loop, one CUDA graph replay per step), CUDA-event timed, max over ranks.
* `e2e`    : the same metric through the reference-shaped Python API (diffusion.p_sample) with the
             step's inputs coming from pinned host memory and the sample read back to the host,
             every step, inside the timed region.
* `roofline`: the dominant kernel class (the 3x3x3 implicit-GEMM convolution) from a profiled pass
             of the same step (CUDA events around every launch on the launching stream).
* `cpu_baseline` / `--impl reference`: the CPU oracle (a port of the reference's PyTorch path; the
             reference itself cannot travel to the GPU box) on a bounded z-slab of the same patch.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

PATCH = (1, 1, 96, 96, 96)
FLOPS_PER_EVAL = 19.68e12  # SURVEY.md section 6 (meta-device trace of the reference)
PATCHES_PER_VOLUME = 18    # scripts/test.py:205-230 tiling of a (110,200,200) volume
C2_FLAGS = dict(
    large_size=96, small_size=96, class_cond=False, learn_sigma=True, num_channels=128, num_res_blocks=2,
    num_heads=4, num_head_channels=64, num_heads_upsample=-1, attention_resolutions="1000", dropout=0.0,
    diffusion_steps=1000, noise_schedule="linear", timestep_respacing="", use_kl=False, predict_xstart=False,
    rescale_timesteps=False, rescale_learned_sigmas=False, use_checkpoint=False, use_scale_shift_norm=True,
    resblock_updown=True, use_fp16=True,
)


def synth_weights(specs, seed=0):
    """Random-init weights of the shipped architecture (there is no checkpoint offline).  Every
    tensor non-zero (the reference's zero_module init would make the network output exactly 0)."""
    g = torch.Generator().manual_seed(seed)
    sd, fan = {}, {}
    for key, shape in specs:
        stem, leaf = key.rsplit(".", 1)
        if stem.endswith(("in_layers.0", "out_layers.0", ".norm")) or stem == "out.0":
            n = torch.randn(shape, generator=g)
            sd[key] = 1 + 0.1 * n if leaf == "weight" else 0.1 * n
        else:
            if leaf == "weight":
                fan[stem] = math.prod(shape[1:])
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan[stem])
    return sd


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._th = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline_sample(sd, threads, z_slab=16, repeats=1):
    """The oracle (CPU port of the reference's fp32 PyTorch path) on a z-slab of the bench patch.
    Returns (evals/s extrapolated to the full 96^3 patch, description)."""
    from oracle import cases
    from oracle.unet import unet_forward
    torch.set_num_threads(threads)
    flags = dict(C2_FLAGS, use_fp16=False)
    cfg = cases.cfg_from_flags(flags)
    g = torch.Generator().manual_seed(3)
    shape = (1, 1, z_slab, PATCH[3], PATCH[4])
    x = torch.randn(shape, generator=g)
    low = torch.rand(shape, generator=g)
    t = torch.tensor([500])
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = unet_forward(cfg, sd, x, t, low)
        best = min(best, time.perf_counter() - t0)
    assert torch.isfinite(out).all()
    scale = PATCH[2] / z_slab
    return 1.0 / (best * scale), (f"oracle fp32 UNet eval on a (1,1,{z_slab},96,96) z-slab of the 96^3 patch "
                                  f"({best:.2f} s, x{scale:.0f} extrapolated to the full patch; the posterior "
                                  f"update is <0.1% of a CPU step)")


def run_reference(args, rank):
    """--impl reference: the CPU implementation of the path on the host cores (oracle port)."""
    if rank != 0:
        return
    from ddpm3d_b200.unet import _Ctx
    from ddpm3d_b200 import script_util as su
    threads = os.cpu_count() or 1
    model, _ = su.sr_create_model_and_diffusion(**dict(C2_FLAGS, use_fp16=False))
    sd = synth_weights(model._specs)
    del model
    vals = []
    desc = ""
    z = 8 if threads < 32 else 16
    for k in range(args.warmup + args.steps):
        v, desc = cpu_baseline_sample(sd, threads, z_slab=z)
        if k >= args.warmup:
            vals.append(v)
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "unet_evals_per_sec", "value": v, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: paper-default 3D UNet, one 96^3 low-dose-conditioned patch, one DDPM step"},
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "volumes_per_min": v * 60.0 / (PATCHES_PER_VOLUME * 1000),
    }
    print(json.dumps(line), flush=True)


def run_c4(args, rank, world, dev):
    """BASELINE.json config 4: one whole-body volume, Z split over the ranks (halo exchange + GroupNorm sums
    over NCCL inside the library).  value = whole-volume UNet evals/s; scaling is strong."""
    import torch.distributed as dist
    from ddpm3d_b200 import script_util as su, slab
    shape = (1, 1, 640, 192, 192)
    if args.shape:
        z, h, w = (int(v) for v in args.shape.split(","))
        shape = (1, 1, z, h, w)
    Z = shape[2]
    model, diffusion = su.sr_create_model_and_diffusion(**C2_FLAGS)
    model.load_state_dict(synth_weights(model._specs))
    model.to(dev)
    model.convert_to_fp16()
    model.eval()
    for o in args.opt:
        model.set_option(o.split("=")[0], int(o.split("=")[1]))
    bounds = slab.slab_bounds(Z, world)
    z0, z1 = bounds[rank], bounds[rank + 1]
    if world > 1:
        model.enable_slab_sharding()
        model.set_slab(z0, Z)
    g = torch.Generator().manual_seed(1234)
    lshape = (1, 1, z1 - z0, shape[3], shape[4])
    low = torch.rand(lshape, generator=g).to(dev)
    x_T = torch.randn(lshape, generator=g).to(dev)
    kw = {"low_res": low}
    K, Wm = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    model._sample_loop(diffusion, x_T, kw, None, 1, True, n_steps=Wm)
    barrier()
    l0 = model.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index) as clocks:
        barrier()
        e0.record()
        out = model._sample_loop(diffusion, x_T, kw, None, 1, True, n_steps=K)
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    launches = model.launch_count() - l0
    assert torch.isfinite(out).all()
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt[0])
    breakdown = {}
    if rank == 0 or world > 1:  # the profiled pass contains collectives: every rank runs it
        model.set_option("profile", 1)
        model._sample_loop(diffusion, x_T, kw, None, 1, True, n_steps=1)
        for kind, t_ms, work in model.profile_read():
            b = breakdown.setdefault(kind, {"ms": 0.0, "work": 0.0, "launches": 0})
            b["ms"] += t_ms
            b["work"] += work
            b["launches"] += 1
        model.set_option("profile", 0)
    if rank == 0:
        vox = shape[2] * shape[3] * shape[4]
        flops = FLOPS_PER_EVAL * vox / 96 ** 3
        line = {
            "metric": "unet_evals_per_sec", "value": K / (ms * 1e-3), "unit": "evals/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"C4: paper-default 3D UNet on ONE {shape[2]}x{shape[3]}x{shape[4]} volume, Z split into "
                                   f"{world} slab(s); one DDPM reverse step per bench step",
                       "parallelism": f"z-slabs x{world}: 1-plane halo per 3x3x3 conv (NCCL send/recv), GroupNorm fp64 sums "
                                      "all-gathered"},
            "tflops_total": flops * K / (ms * 1e-3) / 1e12, "gpu_launches": int(launches), "clocks": clocks.summary(),
            "kernel_breakdown_ms_per_step_rank0": breakdown,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--conv-path", type=int, default=0, help="0 auto, 1 force SIMT, 2 force tcgen05")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shape", default="", help="override Z,H,W (debug only; invalidates the metric)")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (ddpm3d_set_option), repeatable")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"],
                    help="c2 (default, the driver's metric): one 96^3 patch per GPU; c4: ONE 640x192x192 volume "
                         "sharded as z-slabs over the GPUs (strong scaling; extra measurement, BASELINE config 4)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from ddpm3d_b200 import script_util as su

    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    if args.workload == "c4":
        return run_c4(args, rank, world, dev)
    shape = PATCH
    if args.shape:
        z, h, w = (int(v) for v in args.shape.split(","))
        shape = (1, 1, z, h, w)
    flops_per_eval = FLOPS_PER_EVAL * (shape[2] * shape[3] * shape[4]) / (96 ** 3)

    model, diffusion = su.sr_create_model_and_diffusion(**C2_FLAGS)
    sd = synth_weights(model._specs)
    model.load_state_dict(sd)
    model.to(dev)
    model.convert_to_fp16()
    model.eval()
    if args.conv_path:
        model.set_option("conv_path", args.conv_path)
    for o in args.opt:
        model.set_option(o.split("=")[0], int(o.split("=")[1]))

    # independent patches per rank (scripts/test.py:235-246): weak scaling, no data-path collective
    g = torch.Generator().manual_seed(1234 + rank)
    low_h = torch.rand(shape, generator=g).pin_memory()
    g2 = torch.Generator().manual_seed(10)
    xT_h = torch.randn(shape, generator=g2).pin_memory()
    low = low_h.to(dev)
    x_T = xT_h.to(dev)
    kw = {"low_res": low}
    K, Wm = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (captures the graphs, allocates the workspace) ----------------------------------
    model._sample_loop(diffusion, x_T, kw, None, 1, True, n_steps=Wm)
    barrier()

    # ---- timed: K device-resident steps ---------------------------------------------------------
    l0 = model.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        out = model._sample_loop(diffusion, x_T, kw, None, 1, True, n_steps=K)
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    launches = model.launch_count() - l0
    assert torch.isfinite(out).all()

    # ---- e2e: reference-shaped API, host buffers in and out every step -----------------------------
    out_h = torch.empty(shape).pin_memory()
    nz = torch.empty(shape, device=dev)
    t_top = diffusion.num_timesteps - 1

    t_dev = [torch.tensor([t_top - (i % (t_top + 1))], device=dev) for i in range(max(K, Wm))]

    def e2e_step(i):  # the public call: diffusion.p_sample(model, x, t, model_kwargs=..., noise=...)
        x = xT_h.to(dev, non_blocking=True)
        lr = low_h.to(dev, non_blocking=True)
        nz.normal_()
        o = diffusion.p_sample(model, x, t_dev[i], clip_denoised=True, model_kwargs={"low_res": lr}, noise=nz)
        out_h.copy_(o["sample"], non_blocking=True)

    for i in range(Wm):
        e2e_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(K):
        e2e_step(i)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    if world > 1:
        tt = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(tt[0]), float(tt[1])

    # ---- roofline: profiled pass of the same step (rank 0) ----------------------------------------
    roofline = None
    breakdown = {}
    if rank == 0:
        model.set_option("profile", 1)
        model._sample_loop(diffusion, x_T, kw, None, 1, True, n_steps=2)
        recs = model.profile_read()
        model.set_option("profile", 0)
        for kind, t_ms, work in recs:
            b = breakdown.setdefault(kind, {"ms": 0.0, "work": 0.0, "launches": 0})
            b["ms"] += t_ms / 2
            b["work"] += work / 2
            b["launches"] += 0.5
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0  # kernel timed inside a long step
        which = "measured (sustained)" if "bf16_tflops_sustained" in peaks else "fallback"
        dom = "conv_tcgen05" if breakdown.get("conv_tcgen05", {}).get("ms", 0) > 0 else "conv_simt"
        d = breakdown[dom]
        achieved = d["work"] / (d["ms"] * 1e-3) / 1e12
        # dram__bytes_read.sum + dram__bytes_write.sum of the dominant launch (96^3 128->128 convolution, 782.8 GFLOP,
        # algorithmic 226 MB in + 226 MB out + 0.9 MB weights) from the ncu --set full capture in
        # profiles/r1k_conv_tc_ncu_full.txt; only meaningful for the default shape
        traffic = 227427584 + 183572480 if (dom == "conv_tcgen05" and shape == PATCH) else None
        roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved / peak_tf, "traffic": traffic,
                    "traffic_note": "DRAM bytes of the dominant launch (96^3 128->128 conv) from profiles/r1k_conv_tc_ncu_full.txt; "
                                    "achieved/frac are over all 70 conv launches of the step",
                    "peak_source": which,
                    "share_of_step": d["ms"] / sum(b["ms"] for b in breakdown.values()),
                    "launches_per_step": d["launches"]}
        hbm = peaks.get("hbm_gbs", 6650.0)
        for k, b in breakdown.items():
            if k.startswith("conv") or k == "attention":
                b["tflops"] = b["work"] / (b["ms"] * 1e-3) / 1e12 if b["ms"] > 0 else 0.0
            elif b["work"] > 0:
                b["gbps"] = b["work"] / (b["ms"] * 1e-3) / 1e9
                b["frac_of_hbm"] = b["gbps"] / hbm

    # ---- CPU baseline (rank 0, N=1 only) ----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sd32 = {k: v.float() for k, v in sd.items()}
        v, desc = cpu_baseline_sample(sd32, threads, z_slab=8 if threads < 32 else 16)
        cpu = {"value": v, "unit": "evals/s", "cores": threads, "kind": "port", "sample": desc}

    if rank == 0:
        n = world
        value = n * K / (ms * 1e-3)
        e2e_v = n * K / (ms_e2e * 1e-3)
        nbytes = 4 * shape[2] * shape[3] * shape[4]
        line = {
            "metric": "unet_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": n, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "C2: paper-default 3D UNet (128ch, 2 res blocks, mult 1-1-2-3-4), one "
                                   f"{shape[2]}x{shape[3]}x{shape[4]} low-dose-conditioned patch per GPU, one DDPM reverse step "
                                   "(UNet eval + posterior update) per bench step",
                       "l2": "per-step working set (>1 GB of activations + 414 MB of weights) exceeds the 126 MB L2",
                       "weights": "random-init, all tensors non-zero", "parallelism": f"dp{n} (independent patches)"},
            "volumes_per_min": value * 60.0 / (PATCHES_PER_VOLUME * 1000),
            "tflops_per_gpu": flops_per_eval * K / (ms * 1e-3) / 1e12,
            "e2e": {"value": e2e_v, "unit": "evals/s", "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": nbytes},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "kernel_breakdown_ms_per_step": breakdown,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
