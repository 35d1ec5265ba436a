"""A/B of the strip variant on small planes ("strip" level 2) against the brick / stream-K kernel (path 5) for the 12^2 / 6^2
layers of the shipped network, stand-alone (k_conv3d entry point, no folded skip sources)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from ddpm3d_b200 import _native as N
dev = torch.device("cuda", 0)
L = N.lib()


def run(path, Cin, Cout, Z, H, W, reps=30, res=False):
    torch.manual_seed(Cin + Cout + H)
    x = torch.randn((1, Z, H, W, Cin), device=dev).bfloat16()
    w = (torch.randn((Cout, 27 * Cin), device=dev) * 0.02).bfloat16()
    b = torch.zeros(Cout, device=dev)
    r = torch.randn((1, Z, H, W, Cout), device=dev).bfloat16() if res else None
    out = torch.empty((1, Z, H, W, Cout), device=dev, dtype=torch.bfloat16)
    s = N.current_stream_ptr(dev)

    def call():
        N.check(L.ddpm3d_k_conv3d(N.BF16, path, N.ptr(x), N.ptr(w), N.ptr(b), N.ptr(r), N.ptr(out), 1, Z, H, W, Cin, Cout, 27, 1, s))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000.0, out


print(f"DDPM3D_STRIP_EFF={os.environ.get('DDPM3D_STRIP_EFF', '(default 60)')}")
SHAPES_CLASSIC = [(128, 128, 96, 96, 96), (128, 128, 96, 48, 48), (256, 256, 96, 24, 24)]
if len(sys.argv) > 1 and sys.argv[1] == "classic":  # the large-layer shapes only, strip variant: `DDPM3D_LIB=... python ... classic`
    for (Cin, Cout, Z, H, W) in SHAPES_CLASSIC * 2:
        fl = 2.0 * Z * H * W * Cout * 27 * Cin
        t2, _ = run(2, Cin, Cout, Z, H, W)
        print(f"{os.environ.get('DDPM3D_LIB', 'libddpm3d.so')[-18:]:18s} Cin {Cin:4d} Cout {Cout:4d} {Z:4d}x{H}x{W}: {t2:7.1f} us ({fl / t2 / 1e6:6.0f} TF/s)", flush=True)
    sys.exit(0)
for (Cin, Cout, Z, H, W) in [(256, 256, 96, 12, 12), (256, 384, 96, 12, 12), (384, 384, 96, 12, 12), (768, 384, 96, 12, 12),
                             (640, 384, 96, 12, 12), (384, 512, 96, 12, 12), (512, 512, 96, 6, 6), (256, 256, 96, 24, 24),
                             (128, 128, 96, 48, 48), (256, 128, 96, 48, 48), (128, 128, 96, 96, 96), (384, 384, 160, 24, 24), (512, 512, 160, 12, 12)]:
    fl = 2.0 * Z * H * W * Cout * 27 * Cin
    t2, o2 = run(2, Cin, Cout, Z, H, W)
    t7, o7 = run(7, Cin, Cout, Z, H, W)
    t5, o5 = run(5, Cin, Cout, Z, H, W)
    d = float((o2.float() - o5.float()).abs().max() / o5.float().abs().max())
    print(f"Cin {Cin:4d} Cout {Cout:4d} {Z:4d}x{H}x{W}: strip, up to 8 weight stages {t2:7.1f} us ({fl / t2 / 1e6:6.0f} TF/s)   4 stages {t7:7.1f} us "
          f"({fl / t7 / 1e6:6.0f} TF/s)   strip off {t5:7.1f} us ({fl / t5 / 1e6:6.0f} TF/s)   max-rel diff {d:.1e}", flush=True)
