"""Whole-volume denoising around the sampling loop: patch tiling, rank-strided distribution, gather and
Hann-window overlap-add (scripts/test.py:36-183, 185-262, 280-299).  The per-voxel work (patch extraction,
weighted accumulation, normalisation) runs in libddpm3d kernels on the device; this module holds the
integer tiling arithmetic and the torch.distributed plumbing.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from . import dist_util


def calculate_xy_starts(dim_size, patch_size, num_patches=3):
    """scripts/test.py:280-291 (`_calculate_xy_starts_fixed`)."""
    if dim_size == 200 and patch_size == 96 and num_patches == 3:
        return [0, 52, 104]
    if num_patches == 1:
        return [0]
    step = (dim_size - patch_size) / (num_patches - 1)
    starts = [int(i * step) for i in range(num_patches)]
    starts[-1] = min(starts[-1], dim_size - patch_size)
    return starts


def calculate_z_starts(dim_size, patch_size):
    """scripts/test.py:293-299 (`_calculate_z_starts_with_overlap`)."""
    if dim_size <= patch_size:
        return [0]
    return [0, dim_size - patch_size]


def patch_grid(D, H, W, resolution, num_xy=3):
    """Patch origins (z0, h0, w0) in the reference's enumeration order: x (H) outermost, then y (W), then z
    (scripts/test.py:205-218 and :109-111)."""
    return [(z0, x0, y0)
            for x0 in calculate_xy_starts(H, resolution, num_xy)
            for y0 in calculate_xy_starts(W, resolution, num_xy)
            for z0 in calculate_z_starts(D, resolution)]


def hann_window(size):
    """scripts/test.py:248-262: the separable factor and the normaliser of the 3-D window
    outer(outer(h, h).flatten(), h) / max -- max of a product of non-negative factors is (m*m)*m."""
    h = np.hanning(size).astype(np.float64)
    m = float(h.max())
    return h, (m * m) * m


def gather_patches(local, n_patches, device=None):
    """All ranks' samples -> list ordered by patch index on every rank (scripts/test.py:74-80 all_gathers one
    patch per rank per round; that deadlocks when n_patches % world != 0 -- here short ranks pad the last round).
    `local`: {patch index: tensor}; all tensors share one shape."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local[i] for i in range(n_patches)]
    world, rank = dist.get_world_size(), dist.get_rank()
    rounds = (n_patches + world - 1) // world
    proto = next(iter(local.values())) if local else None
    # every rank needs the patch shape even if it owns no patch
    shapes = [None] * world
    dist.all_gather_object(shapes, tuple(proto.shape) if proto is not None else None)
    pshape = next(s for s in shapes if s is not None)
    if device is None:
        device = proto.device if proto is not None else torch.device("cpu")
    out = [None] * n_patches
    for r in range(rounds):
        idx = r * world + rank
        mine = local[idx].contiguous() if idx < n_patches else torch.zeros(pshape, device=device)
        bufs = [torch.empty(pshape, device=device, dtype=mine.dtype) for _ in range(world)]
        dist.all_gather(bufs, mine)
        for k in range(world):
            if r * world + k < n_patches:
                out[r * world + k] = bufs[k]
    return out


def extract_patch(vol_dev, origin, resolution):
    """Device (D,H,W) fp32 -> (1,1,P,P,P) zero-padded patch in (Z,H,W) order (scripts/test.py:219-230, 243-246)."""
    import torch
    D, H, W = vol_dev.shape
    z0, h0, w0 = origin
    out = torch.empty((1, 1, resolution, resolution, resolution), device=vol_dev.device, dtype=torch.float32)
    with torch.cuda.device(vol_dev.device):
        N.check(N.lib().ddpm3d_k_extract_patch(N.ptr(vol_dev), D, H, W, z0, h0, w0, resolution, N.ptr(out),
                                               N.current_stream_ptr(vol_dev.device)))
    return out


def hann_blend(patches, origins, vol_shape, resolution, device):
    """Hann-weighted overlap-add of (Z,H,W)-ordered patches into the reference's (H,W,Z) result
    (scripts/test.py:91-139).  Accumulation order = patch order, arithmetic = numpy's -> bit-identical."""
    import torch
    D, H, W = vol_shape
    arr = torch.zeros((H, W, D), device=device, dtype=torch.float32)
    wsum = torch.zeros_like(arr)
    h, hmax = hann_window(resolution)
    win = torch.from_numpy(h).to(device)
    L = N.lib()
    with torch.cuda.device(device):
        s = N.current_stream_ptr(device)
        for p, (z0, h0, w0) in zip(patches, origins):
            pd = p.to(device, torch.float32).contiguous()
            N.check(L.ddpm3d_k_hann_accumulate(N.ptr(pd), N.ptr(win), C.c_double(hmax), resolution, D, H, W, z0, h0, w0,
                                               N.ptr(arr), N.ptr(wsum), s))
        N.check(L.ddpm3d_k_hann_finalize(N.ptr(arr), N.ptr(wsum), arr.numel(), s))
    return arr, wsum


def denoise_volume(model, diffusion, volume, resolution=96, clip_denoised=True, seed=10, sample_fn=None,
                   log=lambda *_: None, **loop_kwargs):
    """scripts/test.py main(): (D,H,W) low-dose volume -> denoised (H,W,Z) volume (device tensor on rank 0's GPU;
    every rank returns it).  Patches are rank-strided over the process group like scripts/test.py:235-246;
    torch's CUDA generator is seeded with `seed` on every rank like scripts/test.py:45-48."""
    import torch
    dev = next(model.parameters()).device
    vol = torch.as_tensor(np.asarray(volume, dtype=np.float32) if not torch.is_tensor(volume) else volume)
    if vol.dim() == 4 and vol.shape[0] == 1:
        vol = vol[0]
    assert vol.dim() == 3, "volume must be (D, H, W)"
    vol = vol.to(dev, torch.float32).contiguous()
    D, H, W = vol.shape
    origins = patch_grid(D, H, W, resolution)
    mine = dist_util.patch_indices(len(origins))
    log(f"Total patches to process: {len(origins)} (this rank: {len(mine)})")
    if seed is not None:
        torch.cuda.manual_seed_all(seed) if dev.type == "cuda" else torch.manual_seed(seed)
    if sample_fn is None:
        def sample_fn(low_res):
            shape = tuple(low_res.shape)
            noise = torch.randn(*shape, device=dev)
            return diffusion.p_sample_loop(model, shape, noise, clip_denoised=clip_denoised,
                                           model_kwargs={"low_res": low_res}, **loop_kwargs)
    local = {}
    for i in mine:
        low = extract_patch(vol, origins[i], resolution)
        local[i] = sample_fn(low)[0, 0].clone()  # (P,P,P) in (Z,H,W) order
        log(f"Processed patch {i}")
    patches = gather_patches(local, len(origins), device=dev)
    arr, _ = hann_blend(patches, origins, (D, H, W), resolution, dev)
    return arr
