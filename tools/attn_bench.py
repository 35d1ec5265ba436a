"""Times the attention core (tcgen05 vs CUDA-core kernel) at the sizes --attention_resolutions would enable."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ddpm3d_b200 import _native as N
dev = torch.device("cuda", 0)
L = N.lib()
for (B, T, C, heads) in ((1, 3456, 512, 8), (1, 13824, 384, 6)):
    qkv = torch.randn((B, T, 3 * C), device=dev).bfloat16()
    out = torch.empty((B, T, C), device=dev, dtype=torch.bfloat16)
    s = N.current_stream_ptr(dev)
    for path, name in ((0, "tcgen05"), (0x100, "cuda-core")):
        def call():
            N.check(L.ddpm3d_k_attention(N.BF16, N.ptr(qkv), N.ptr(out), B, T, C, heads, path, s))
        reps = 10 if path == 0 else 2
        call(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 4.0 * B * T * T * C
        print(f"T={T} C={C} heads={heads} {name:10s}: {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s (algorithmic 4*T^2*C)", flush=True)
