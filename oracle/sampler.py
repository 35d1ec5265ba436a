"""Oracle: the DDPM ancestral sampler (p_mean_variance / p_sample / p_sample_loop).

TEST INFRASTRUCTURE (see oracle/__init__.py).  fp32 torch-CPU elementwise ops in
exactly the reference's order, so that a CUDA kernel using non-fused fp32
mul/add reproduces it to the last bit except for exp().
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

from .schedule import DiffusionTables


def _pick(arr: np.ndarray, t: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """gaussian_diffusion.py:897-910 _extract_into_tensor: fp64 table -> gather -> .float()."""
    v = torch.from_numpy(np.asarray(arr, dtype=np.float64))[t].float()
    return v.reshape(-1, *([1] * (like.dim() - 1))).expand(like.shape)


def model_timesteps(tabs: DiffusionTables, t: torch.Tensor) -> torch.Tensor:
    """respace.py:123-128 _WrappedModel.__call__: index remap (+ optional rescale)."""
    m = torch.tensor(tabs.timestep_map, dtype=t.dtype)[t]
    if tabs.rescale_timesteps:
        m = m.float() * (1000.0 / tabs.original_num_steps)
    return m


def p_mean_variance(tabs: DiffusionTables, model_output: torch.Tensor, x: torch.Tensor,
                    t: torch.Tensor, clip_denoised: bool = True) -> dict:
    """gaussian_diffusion.py:232-326, from the raw model output onward."""
    B, C = x.shape[:2]
    vt = tabs.model_var_type
    if vt in ("learned", "learned_range"):
        assert model_output.shape == (B, 2 * C, *x.shape[2:])
        model_output, var_values = torch.split(model_output, C, dim=1)
        if vt == "learned":
            log_var = var_values
        else:
            min_log = _pick(tabs.posterior_log_variance_clipped, t, x)
            max_log = _pick(np.log(tabs.betas), t, x)
            frac = (var_values + 1) / 2
            log_var = frac * max_log + (1 - frac) * min_log
        var = torch.exp(log_var)
    else:
        if vt == "fixed_large":
            v = np.append(tabs.posterior_variance[1], tabs.betas[1:])
            var, log_var = _pick(v, t, x), _pick(np.log(v), t, x)
        else:
            var = _pick(tabs.posterior_variance, t, x)
            log_var = _pick(tabs.posterior_log_variance_clipped, t, x)

    def proc(z):
        return z.clamp(-1, 1) if clip_denoised else z

    mt = tabs.model_mean_type
    if mt == "previous_x":
        c1, c2 = tabs.posterior_mean_coef1, tabs.posterior_mean_coef2
        x0 = proc(_pick(1.0 / c1, t, x) * model_output - _pick(c2 / c1, t, x) * x)
        mean = model_output
    else:
        if mt == "start_x":
            x0 = proc(model_output)
        else:
            x0 = proc(_pick(tabs.sqrt_recip_alphas_cumprod, t, x) * x
                      - _pick(tabs.sqrt_recipm1_alphas_cumprod, t, x) * model_output)
        mean = _pick(tabs.posterior_mean_coef1, t, x) * x0 + _pick(tabs.posterior_mean_coef2, t, x) * x
    return dict(mean=mean, variance=var, log_variance=log_var, pred_xstart=x0)


def p_sample(tabs: DiffusionTables, model_output, x, t, noise, clip_denoised=True) -> dict:
    """gaussian_diffusion.py:395-439 (cond_fn=None)."""
    out = p_mean_variance(tabs, model_output, x, t, clip_denoised)
    mask = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
    sample = out["mean"] + mask * torch.exp(0.5 * out["log_variance"]) * noise
    return dict(sample=sample, pred_xstart=out["pred_xstart"], **{k: out[k] for k in ("mean", "log_variance")})


def ddim_sample(tabs: DiffusionTables, model_output, x, t, noise, clip_denoised=True, eta=0.0) -> dict:
    """gaussian_diffusion.py:537-585 (cond_fn=None), same fp32 op order."""
    out = p_mean_variance(tabs, model_output, x, t, clip_denoised)
    x0 = out["pred_xstart"]
    eps = (_pick(tabs.sqrt_recip_alphas_cumprod, t, x) * x - x0) / _pick(tabs.sqrt_recipm1_alphas_cumprod, t, x)
    ab = _pick(tabs.alphas_cumprod, t, x)
    abp = _pick(tabs.alphas_cumprod_prev, t, x)
    sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
    mean_pred = x0 * torch.sqrt(abp) + torch.sqrt(1 - abp - sigma ** 2) * eps
    mask = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
    return dict(sample=mean_pred + mask * sigma * noise, pred_xstart=x0)


@torch.no_grad()
def p_sample_loop(tabs: DiffusionTables, model: Callable, x_T: torch.Tensor, noises,
                  clip_denoised: bool = True, on_step: Optional[Callable] = None) -> torch.Tensor:
    """gaussian_diffusion.py:441-535.  `model(x, t_mapped)` returns the raw UNet
    output; `noises[k]` is the tensor consumed by the k-th executed step
    (k = 0 for i = T-1), replacing th.randn_like (gaussian_diffusion.py:430)."""
    img = x_T
    B = x_T.shape[0]
    for k, i in enumerate(range(tabs.num_timesteps - 1, -1, -1)):
        t = torch.tensor([i] * B)
        mo = model(img, model_timesteps(tabs, t))
        out = p_sample(tabs, mo, img, t, noises[k], clip_denoised)
        if on_step is not None:
            on_step(k, i, mo, out)
        img = out["sample"]
    return img
