"""Helpers shared by the -m gpu parity tests (call the CUDA path through the C ABI)."""
import ctypes as C

import numpy as np
import torch

from ddpm3d_b200 import _native as N

DEV = torch.device("cuda", 0)


def stream():
    return N.current_stream_ptr(DEV)


def to_cl(x, dtype=torch.float32):
    """(B,C,Z,H,W) -> contiguous channels-last (B,Z,H,W,C) on the device."""
    return x.permute(0, 2, 3, 4, 1).contiguous().to(DEV, dtype)


def from_cl(x):
    return x.float().permute(0, 4, 1, 2, 3).contiguous().cpu()


def pack_weight(w, dtype=torch.float32):
    """[Cout,Cin,kd,kh,kw] -> [Cout][taps*Cin] with k = tap*Cin + ci (include/ddpm3d.h)."""
    co, ci = w.shape[:2]
    taps = int(np.prod(w.shape[2:]))
    return w.reshape(co, ci, taps).permute(0, 2, 1).reshape(co, taps * ci).contiguous().to(DEV, dtype)


def max_rel(a, b):
    """||a-b||_inf / ||b||_inf per tensor."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


TDT = {N.FP32: torch.float32, N.BF16: torch.bfloat16, N.FP16: torch.float16}
# one output rounding of the element type (plus fp32 accumulation-order noise)
ROUND_TOL = {N.FP32: 2e-5, N.BF16: 6e-3, N.FP16: 8e-4}


def conv3d(dt, path, x_cl, w_packed, bias, res_cl, B, Z, H, W, Cin, Cout, taps=27, stride=1):
    tdt = TDT[dt]
    out = torch.empty((B, Z, H // stride, W // stride, Cout), device=DEV, dtype=tdt)
    N.check(N.lib().ddpm3d_k_conv3d(dt, path, N.ptr(x_cl), N.ptr(w_packed), N.ptr(bias), N.ptr(res_cl), N.ptr(out),
                                    B, Z, H, W, Cin, Cout, taps, stride, stream()))
    torch.cuda.synchronize()
    return out
