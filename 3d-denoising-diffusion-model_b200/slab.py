"""One large whole-body volume sharded as z-slabs over the GPUs of one box (BASELINE.json config 4; SURVEY.md
section 8e.3).  The network never strides Z, so each rank runs the whole sampler on its slab; the library
exchanges one halo plane per 3x3x3 conv and all-gathers GroupNorm sums (csrc/comm.cu).  This module only
partitions, scatters and gathers."""
from __future__ import annotations


def slab_bounds(z_total, world):
    """Contiguous, nearly equal slabs: rank r owns [b[r], b[r+1])."""
    if world < 1 or z_total < world:
        raise ValueError("need at least one plane per rank")
    base, extra = divmod(z_total, world)
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < extra else 0))
    return bounds


def gather_slabs(local, bounds, group=None):
    """All ranks' (B,C,Zl,H,W) slabs -> the full (B,C,Z,H,W) tensor on every rank."""
    import torch
    import torch.distributed as dist
    world = len(bounds) - 1
    if world == 1:
        return local
    zmax = max(bounds[r + 1] - bounds[r] for r in range(world))
    B, Cc, Zl, H, W = local.shape
    pad = torch.zeros((B, Cc, zmax, H, W), device=local.device, dtype=local.dtype)
    pad[:, :, :Zl] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][:, :, :bounds[r + 1] - bounds[r]] for r in range(world)], dim=2)


def sample_volume_slabs(model, diffusion, low_res, noise=None, clip_denoised=True, rng="philox", seed=0, step_noise=None,
                        group=None):
    """p_sample_loop of ONE (B,1,Z,H,W) volume with Z split over the ranks of `group`.  `low_res` / `noise` (x_T) /
    `step_noise` ([T,B,1,Z,H,W]) are full-volume tensors available on every rank; each rank slices its slab.
    With rng="philox" every rank draws its part of one global noise field, so the result equals the
    single-GPU philox run.  model.enable_slab_sharding(group) must have been called."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = next(model.parameters()).device
    Z = low_res.shape[2]
    bounds = slab_bounds(Z, world)
    z0, z1 = bounds[rank], bounds[rank + 1]
    model.set_slab(z0, Z)
    low = low_res[:, :, z0:z1].to(dev).contiguous()
    if noise is None:
        raise ValueError("pass x_T (`noise`) for the whole volume so every rank slices the same field")
    x_T = noise[:, :, z0:z1].to(dev).contiguous()
    sn = step_noise[:, :, :, z0:z1].to(dev).contiguous() if step_noise is not None else None
    out = diffusion.p_sample_loop(model, tuple(low.shape), noise=x_T, clip_denoised=clip_denoised,
                                  model_kwargs={"low_res": low}, rng=rng, seed=seed, step_noise=sn)
    return gather_slabs(out, bounds, group)
