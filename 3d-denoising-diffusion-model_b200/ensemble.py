"""Uncertainty-map ensemble (README.md:44 "uncertainty maps"; BASELINE.json config 5): several stochastic
samples of the same low-dose input, voxel-wise mean and variance reduced on the device.  Samples shard over
ranks with no communication; per-rank Welford partials are merged once at the end in rank order."""
from __future__ import annotations

from . import _native as N
from . import dist_util


class Welford:
    """Running voxel-wise mean / M2 on the device (libddpm3d kernels)."""

    def __init__(self, shape, device):
        import torch
        self.mean = torch.zeros(shape, device=device, dtype=torch.float32)
        self.m2 = torch.zeros(shape, device=device, dtype=torch.float32)
        self.count = 0
        self.device = device

    def update(self, x):
        import torch
        x = x.to(self.device, torch.float32).contiguous()
        assert x.shape == self.mean.shape
        self.count += 1
        with torch.cuda.device(self.device):
            N.check(N.lib().ddpm3d_k_welford_update(N.ptr(self.mean), N.ptr(self.m2), N.ptr(x), self.count,
                                                    x.numel(), N.current_stream_ptr(self.device)))

    def merge(self, mean_b, m2_b, count_b):
        import torch
        if count_b == 0:
            return
        mean_b = mean_b.to(self.device, torch.float32).contiguous()
        m2_b = m2_b.to(self.device, torch.float32).contiguous()
        with torch.cuda.device(self.device):
            N.check(N.lib().ddpm3d_k_welford_merge(N.ptr(self.mean), N.ptr(self.m2), self.count, N.ptr(mean_b),
                                                   N.ptr(m2_b), count_b, mean_b.numel(),
                                                   N.current_stream_ptr(self.device)))
        self.count += count_b

    def variance(self, unbiased=True):
        d = self.count - 1 if unbiased else self.count
        return self.m2 / max(d, 1)


def gather_partials(mean, m2, count):
    """Every rank's (mean, M2, count), in rank order (torch.distributed plumbing)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [(mean, m2, count)]
    world = dist.get_world_size()
    means = [torch.empty_like(mean) for _ in range(world)]
    m2s = [torch.empty_like(m2) for _ in range(world)]
    counts = [None] * world
    dist.all_gather(means, mean.contiguous())
    dist.all_gather(m2s, m2.contiguous())
    dist.all_gather_object(counts, int(count))
    return list(zip(means, m2s, counts))


def ensemble_sample(model, diffusion, low_res, seeds, clip_denoised=True, **loop_kwargs):
    """One sample per seed (seed -> torch CUDA generator, like scripts/test.py:45-48 does with 10), seeds
    rank-strided; returns (mean, variance, n) identical on every rank."""
    import torch
    dev = next(model.parameters()).device
    low_res = low_res.to(dev)
    shape = tuple(low_res.shape)
    acc = Welford(shape, dev)
    for k in dist_util.patch_indices(len(seeds)):
        torch.cuda.manual_seed_all(int(seeds[k]))
        noise = torch.randn(*shape, device=dev)
        acc.update(diffusion.p_sample_loop(model, shape, noise, clip_denoised=clip_denoised,
                                           model_kwargs={"low_res": low_res}, **loop_kwargs))
    total = Welford(shape, dev)
    for mean, m2, count in gather_partials(acc.mean, acc.m2, acc.count):
        total.merge(mean, m2, count)
    return total.mean, total.variance(), total.count
