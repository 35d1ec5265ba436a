"""Oracle helpers: deterministic synthetic weights and inputs.

TEST INFRASTRUCTURE (see oracle/__init__.py).  There is no released checkpoint
in the sandbox, and a freshly constructed reference model outputs exactly 0
(zero_module on every out_layers conv, proj_out and the final conv --
guided_diffusion/unet.py:210-212,294,996; SURVEY.md section 4 "random-init
trap").  So every parameter is drawn non-zero here, with the same
distributions torch's default initialisers use, from a CPU generator so the
bits are identical wherever they are produced.
"""
from __future__ import annotations

import math

import torch

from .unet import UNetConfig, param_specs


def _is_norm(key: str) -> bool:
    stem = key.rsplit(".", 1)[0]
    return stem.endswith(("in_layers.0", "out_layers.0", ".norm")) or stem == "out.0"


def synth_state_dict(cfg: UNetConfig, seed: int = 0) -> dict:
    """fp32 state_dict with the reference's keys/shapes (param_specs)."""
    sd = {}
    fan_in = {}
    for idx, (key, shape) in enumerate(param_specs(cfg)):
        g = torch.Generator().manual_seed(seed * 1000003 + idx)
        stem, leaf = key.rsplit(".", 1)
        if _is_norm(key):
            n = torch.randn(shape, generator=g)
            sd[key] = (1.0 + 0.1 * n) if leaf == "weight" else 0.1 * n
        elif key == "label_emb.weight":
            sd[key] = torch.randn(shape, generator=g)
        else:
            if leaf == "weight":
                fan_in[stem] = math.prod(shape[1:])
            b = 1.0 / math.sqrt(fan_in[stem])
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * b
    return sd


def synth_inputs(shape, n_steps: int, seed_cond: int = 1234, seed_noise: int = 10, phantom: bool = False):
    """low_res (seed 1234), x_T and the per-step noise list (seed 10, the seed
    scripts/test.py:45-48 fixes).  All CPU fp32; callers copy to the device so
    both paths see identical bits."""
    g = torch.Generator().manual_seed(seed_cond)
    if phantom:
        B, _, Z, H, W = shape
        zz, hh, ww = torch.meshgrid(
            torch.linspace(0, 1, Z), torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
        vol = torch.zeros(Z, H, W)
        for _ in range(8):
            c = torch.rand(3, generator=g)
            s = 0.05 + 0.2 * torch.rand(1, generator=g)
            a = torch.rand(1, generator=g)
            vol += a * torch.exp(-((zz - c[0]) ** 2 + (hh - c[1]) ** 2 + (ww - c[2]) ** 2) / (2 * s * s))
        vol = vol / vol.max()
        vol = (vol + torch.sqrt(vol / 20) * torch.randn(vol.shape, generator=g)).clamp_min(0)
        low = vol[None, None].expand(B, 1, Z, H, W).contiguous()
    else:
        low = torch.rand(shape, generator=g)
    g2 = torch.Generator().manual_seed(seed_noise)
    x_T = torch.randn(shape, generator=g2)
    noises = [torch.randn(shape, generator=g2) for _ in range(n_steps)]
    return low, x_T, noises
