"""Which descriptor encoding lets tcgen05 read an A operand that starts at an arbitrary row of a SWIZZLE_128B tile?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ddpm3d_b200 import _native as N
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
A = torch.randn((256, 64), generator=g).bfloat16().to(dev)
I = torch.eye(64).bfloat16().to(dev)
out = torch.empty((128, 64), device=dev)
for mode in (0, 1):
    for shift in (0, 1, 2, 3, 5, 7, 8, 9, 16, 21, 64, 100, 127):
        N.check(N.lib().ddpm3d_k_probe_rowshift(N.ptr(A), 256, N.ptr(I), shift, mode, N.ptr(out), N.current_stream_ptr(dev)))
        torch.cuda.synchronize()
        want = A[shift:shift + 128].float()
        ok = torch.equal(out, want)
        extra = ""
        if not ok:
            # which source row did each output row get?
            src = [(int((A.float() == out[r]).all(dim=1).nonzero()[0]) if (A.float() == out[r]).all(dim=1).any() else -1) for r in range(16)]
            extra = f" first rows read from {src}"
        print(f"mode {mode} shift {shift:3d}: {'OK' if ok else 'MISMATCH'}{extra}", flush=True)
