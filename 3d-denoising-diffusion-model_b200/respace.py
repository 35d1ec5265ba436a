"""Timestep respacing (guided_diffusion/respace.py).  Integer / fp64 host arithmetic that must be
bit-exact with the reference, so it stays in Python (Python's round() on a float accumulator is
part of the contract, respace.py:53-57)."""
from __future__ import annotations

import numpy as np

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """respace.py:7-60: the set of original steps a respaced process keeps.
    `section_counts`: "N", "a,b,c", "ddimN" or a list of ints."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            target = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == target:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    per, extra = divmod(num_timesteps, len(section_counts))
    kept = []
    start = 0
    for k, want in enumerate(section_counts):
        size = per + (1 if k < extra else 0)
        if size < want:
            raise ValueError(f"cannot divide section of {size} steps into {want}")
        frac_stride = 1 if want <= 1 else (size - 1) / (want - 1)
        pos = 0.0
        for _ in range(want):
            kept.append(start + round(pos))
            pos += frac_stride
        start += size
    return set(kept)


class SpacedDiffusion(GaussianDiffusion):
    """respace.py:63-113: a diffusion process over a subset of the base steps."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        base_betas = np.array(kwargs["betas"], dtype=np.float64)
        self.original_num_steps = len(base_betas)
        base_acp = np.cumprod(1.0 - base_betas, axis=0)
        self.timestep_map = []
        prev = 1.0
        betas = []
        for i, acp in enumerate(base_acp):
            if i in self.use_timesteps:
                betas.append(1 - acp / prev)
                prev = acp
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(betas)
        super().__init__(**kwargs)

    def _map_timesteps(self, t):
        """respace.py:116-128 _WrappedModel.__call__."""
        import torch
        m = torch.tensor(self.timestep_map, device=t.device, dtype=t.dtype)[t]
        if self.rescale_timesteps:
            m = m.float() * (1000.0 / self.original_num_steps)
        return m

    def _scale_timesteps(self, t):
        # respace.py:111-113: scaling is done by the wrapped model
        return t
