"""Times ddpm3d_k_conv3d (tcgen05 path) on synthetic layers: separates per-kernel, per-tile and per-k-step cost."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from ddpm3d_b200 import _native as N
dev = torch.device("cuda", 0)
L = N.lib()
def run(Cin, Cout, Z, H, W, reps=20, res=False):
    x = torch.randn((1, Z, H, W, Cin), device=dev).bfloat16()
    w = (torch.randn((Cout, 27 * Cin), device=dev) * 0.02).bfloat16()
    b = torch.zeros(Cout, device=dev)
    r = torch.randn((1, Z, H, W, Cout), device=dev).bfloat16() if res else None
    out = torch.empty((1, Z, H, W, Cout), device=dev, dtype=torch.bfloat16)
    s = N.current_stream_ptr(dev)
    def call():
        N.check(L.ddpm3d_k_conv3d(N.BF16, 2, N.ptr(x), N.ptr(w), N.ptr(b), N.ptr(r), N.ptr(out), 1, Z, H, W, Cin, Cout, 27, 1, s))
    for _ in range(3): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * Z * H * W * Cout * 27 * Cin
    print(f"Cin {Cin:4d} Cout {Cout:4d} {Z:4d}x{H}x{W} res={int(res)}: {ms*1000:8.1f} us  {fl/ms/1e9:7.0f} TFLOP/s", flush=True)
print("# 24^2 level, Cout=256 (BN=256 tiles): vary waves (Z) and k-steps (Cin)")
for Z in (33, 66, 96, 192, 384):
    for Cin in (256, 512):
        run(Cin, 256, Z, 24, 24)
print("# 96^2 level, Cout=128 (two bricks per CTA)")
for Z in (12, 24, 48, 96):
    for Cin in (128, 256):
        run(Cin, 128, Z, 96, 96)
print("# 48^2, 12^2, 6^2")
run(128, 128, 96, 48, 48); run(256, 128, 96, 48, 48); run(128, 128, 96, 48, 48, res=True)
run(384, 384, 96, 12, 12); run(768, 384, 96, 12, 12); run(512, 512, 96, 6, 6); run(1024, 512, 96, 6, 6)
