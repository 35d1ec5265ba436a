"""Oracle: patch tiling and Hann-window overlap-add (scripts/test.py), numpy on the CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The tiling helpers are pinned against the reference's own
functions by tests/golden/volume.npz (oracle/make_golden.py imports scripts/test.py with its missing
third-party imports stubbed); the blend restates the body of main() (scripts/test.py:91-139), which the
reference does not expose as a function.
"""
from __future__ import annotations

import numpy as np


def xy_starts(dim_size, patch_size, num_patches=3):
    """scripts/test.py:280-291."""
    if dim_size == 200 and patch_size == 96 and num_patches == 3:
        return [0, 52, 104]
    if num_patches == 1:
        return [0]
    step = (dim_size - patch_size) / (num_patches - 1)
    out = [int(k * step) for k in range(num_patches)]
    out[-1] = min(out[-1], dim_size - patch_size)
    return out


def z_starts(dim_size, patch_size):
    """scripts/test.py:293-299."""
    return [0] if dim_size <= patch_size else [0, dim_size - patch_size]


def hann3d(size):
    """scripts/test.py:248-262."""
    h = np.hanning(size)
    w = np.outer(np.outer(h, h).flatten(), h).reshape(size, size, size)
    return w / w.max()


def make_patches(vol, resolution):
    """scripts/test.py:192-233: (D,H,W) -> list of zero-padded (Z,H,W) patches, x/y/z loop order."""
    D, H, W = vol.shape
    out, origins = [], []
    for x0 in xy_starts(H, resolution):
        for y0 in xy_starts(W, resolution):
            for z0 in z_starts(D, resolution):
                p = vol[z0:min(z0 + resolution, D), x0:min(x0 + resolution, H), y0:min(y0 + resolution, W)]
                pad = np.zeros((resolution,) * 3, dtype=np.float32)
                pad[:p.shape[0], :p.shape[1], :p.shape[2]] = p
                out.append(pad)
                origins.append((z0, x0, y0))
    return out, origins


def blend(patches_zhw, vol_shape, resolution):
    """scripts/test.py:72,91-139: patches arrive as (Z,H,W), are permuted to (H,W,Z), Hann-weighted and
    accumulated into float32 arrays (numpy promotes the fp32*fp64 product to fp64 and stores fp32)."""
    D, H, W = vol_shape
    arr = np.zeros((H, W, D), dtype=np.float32)
    wsum = np.zeros_like(arr)
    win = hann3d(resolution)
    k = 0
    for x0 in xy_starts(H, resolution):
        for y0 in xy_starts(W, resolution):
            for z0 in z_starts(D, resolution):
                patch = np.transpose(patches_zhw[k], (1, 2, 0))
                x1, y1, z1 = min(x0 + resolution, H), min(y0 + resolution, W), min(z0 + resolution, D)
                hx, wy, dz = x1 - x0, y1 - y0, z1 - z0
                arr[x0:x1, y0:y1, z0:z1] += patch[:hx, :wy, :dz] * win[:hx, :wy, :dz]
                wsum[x0:x1, y0:y1, z0:z1] += win[:hx, :wy, :dz]
                k += 1
    # scripts/test.py:139 calls np.divide(arr, wsum, where=wsum > 0) WITHOUT `out=`: voxels whose total Hann
    # weight is 0 (the outer faces of the volume, np.hanning(P)[0] == 0) come back as uninitialised memory in
    # the reference.  They are pinned to the accumulated value (0) here and in the CUDA path.
    out = arr.copy()
    np.divide(arr, wsum, out=out, where=wsum > 0)
    return out, wsum
