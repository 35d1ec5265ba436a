"""HBM-bound kernels stand-alone at sizes that fill the GPU: the p_sample update on C4's
23.6 M voxels, GroupNorm32 + SiLU (stats + finalize + apply) plain / pooled / upsampled on a 96^3 x 128 bf16 tensor.
Prints algorithmic GB/s (CUDA events, 20 repetitions after 3 warm-ups) against the measured copy peak.
    python tools/elementwise_bench.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ddpm3d_b200 import _native as N, script_util as su  # noqa: E402
from ddpm3d_b200.unet import sampler_only_context  # noqa: E402

dev = torch.device("cuda", 0)
peak = 6448.4
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
stream = lambda: N.current_stream_ptr(dev)  # noqa: E731


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rows = []
# ---- p_sample update: x, eps, v, noise read; sample, pred_xstart written = 24 B/voxel ---------------------------
diffusion = su.create_gaussian_diffusion(steps=1000, learn_sigma=True, timestep_respacing="100")
ctx = sampler_only_context(diffusion, dev)
n = 640 * 192 * 192
x = torch.randn((1, 1, n), device=dev)
mo = torch.randn((1, 2, n), device=dev)
nz = torch.randn((1, 1, n), device=dev)
ti = torch.tensor([50], dtype=torch.int32, device=dev)
smp, x0 = torch.empty_like(x), torch.empty_like(x)
ms = timeit(lambda: N.check(N.lib().ddpm3d_p_sample_update(ctx, N.ptr(x), N.ptr(mo), N.ptr(nz), N.ptr(ti), 1, N.ptr(smp), N.ptr(x0),
                                                           None, None, 1, 1, n, stream())))
rows.append(("p_sample update, 640x192x192 (learned range, clip)", 24.0 * n, ms))
del x, mo, nz, smp, x0

# ---- GroupNorm32 + SiLU, bf16, 96^3 x 128: stats read + apply read + write ----------------------------------------
B, Z, H, W, C = 1, 96, 96, 96, 128
xin = torch.randn((B, Z, H, W, C), device=dev).bfloat16()
gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
for name, mode, oscale in (("GroupNorm+SiLU", 0, 1.0), ("GroupNorm+SiLU+avgpool(1,2,2)", 1, 0.25), ("GroupNorm+SiLU+nearest x2", 2, 4.0)):
    Ho, Wo = (H // 2, W // 2) if mode == 1 else ((2 * H, 2 * W) if mode == 2 else (H, W))
    out = torch.empty((B, Z, Ho, Wo, C), device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: N.check(N.lib().ddpm3d_k_groupnorm(N.BF16, N.ptr(xin), N.ptr(gam), N.ptr(bet), None, 1, mode, N.ptr(out),
                                                           B, Z, H, W, C, stream())))
    rows.append((f"{name}, 96^3 x 128 bf16 (3 kernels)", xin.numel() * 2 * (2 + oscale), ms))
    del out

print(f"# algorithmic bytes / time, stand-alone, peak = {peak:.0f} GB/s (measured copy)")
for name, nbytes, ms in rows:
    gbs = nbytes / ms / 1e6
    print(f"{name:58s} {nbytes / 1e6:9.1f} MB {ms * 1e3:8.1f} us {gbs:8.0f} GB/s {100 * gbs / peak:5.1f} %")
