"""Uncertainty-map ensemble (README.md:44 "uncertainty maps"; BASELINE.json config 5): several stochastic
samples of the same low-dose input, voxel-wise mean and variance reduced on the device.  Samples shard over
ranks with no communication; per-rank Welford partials are merged once at the end along a fixed binary tree."""
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


def reduce_partials(acc, group=None):
    """Merges every rank's Welford partial into rank 0's `acc` along a fixed binary tree (log2(world) rounds of one
    point-to-point (mean, M2) transfer each; the merge order depends only on the world size, so the result is
    deterministic).  Returns True on the rank that holds the total."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return True
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    g = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    step = 1
    while step < world:
        if rank % (2 * step) == step:  # sender: hands its partial to rank - step and is done
            cnt = torch.tensor([acc.count], device=acc.device, dtype=torch.int64)
            dist.send(cnt, g(rank - step), group=group)
            if acc.count:
                dist.send(acc.mean.contiguous(), g(rank - step), group=group)
                dist.send(acc.m2.contiguous(), g(rank - step), group=group)
            return False
        if rank % (2 * step) == 0 and rank + step < world:
            cnt = torch.zeros(1, device=acc.device, dtype=torch.int64)
            dist.recv(cnt, g(rank + step), group=group)
            n_b = int(cnt.item())
            if n_b:
                mean_b, m2_b = torch.empty_like(acc.mean), torch.empty_like(acc.m2)
                dist.recv(mean_b, g(rank + step), group=group)
                dist.recv(m2_b, g(rank + step), group=group)
                acc.merge(mean_b, m2_b, n_b)
        step *= 2
    return rank == 0


def ensemble_sample(model, diffusion, low_res, seeds, clip_denoised=True, all_ranks=True, **loop_kwargs):
    """One sample per seed (seed -> torch CUDA generator, like scripts/test.py:45-48 does with 10), seeds
    rank-strided; the per-rank partials are reduced to rank 0 (tree) and, with `all_ranks`, the result is broadcast.
    Returns (mean, variance, n): valid on rank 0, and on every rank when `all_ranks`."""
    import torch
    dev = next(model.parameters()).device
    low_res = low_res.to(dev)
    shape = tuple(low_res.shape)
    acc = Welford(shape, dev)
    for k in dist_util.patch_indices(len(seeds)):
        torch.cuda.manual_seed_all(int(seeds[k]))
        noise = torch.randn(*shape, device=dev)
        kw = dict(loop_kwargs)
        if kw.get("rng") == "philox":  # device-resident loop: the per-step noise is drawn in-kernel from the sample's seed
            kw["seed"] = int(seeds[k])
        acc.update(diffusion.p_sample_loop(model, shape, noise, clip_denoised=clip_denoised,
                                           model_kwargs={"low_res": low_res}, **kw))
    reduce_partials(acc)
    mean, var, n = acc.mean, acc.variance(), acc.count
    if all_ranks:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            cnt = torch.tensor([n], device=dev, dtype=torch.int64)
            dist.broadcast(cnt, 0)
            n = int(cnt.item())
            var = acc.m2 / max(n - 1, 1) if dist.get_rank() == 0 else torch.empty_like(mean)
            dist.broadcast(mean, 0)
            dist.broadcast(var, 0)
    return mean, var, n
