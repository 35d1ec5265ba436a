"""CPU: tiling / Hann helpers (oracle pinned to the reference's functions, host mirror equal to the oracle)
and the world_size-2 gloo tests of the N>1 host logic (rank-strided patches, gather with n % world != 0,
ensemble partial tree reduction)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ddpm3d_b200 import dist_util, volume
from oracle import volume as ov
from oracle.cases import HANN_SIZES, VOLUME_DIMS, VOLUME_Z

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_tiling_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "volume.npz"))
    for i, (dim, p, n) in enumerate(VOLUME_DIMS):
        assert ov.xy_starts(dim, p, n) == g[f"xy{i}"].tolist()
        assert volume.calculate_xy_starts(dim, p, n) == g[f"xy{i}"].tolist()
    for i, (dim, p) in enumerate(VOLUME_Z):
        assert ov.z_starts(dim, p) == g[f"z{i}"].tolist()
        assert volume.calculate_z_starts(dim, p) == g[f"z{i}"].tolist()
    for s in HANN_SIZES:
        w = ov.hann3d(s).reshape(-1)[::7 if s == 96 else 1]
        assert np.array_equal(w.view(np.int64), g[f"hann{s}"].view(np.int64))
        h, m = volume.hann_window(s)
        full = (np.multiply.outer(np.multiply.outer(h, h), h) / m).reshape(-1)[::7 if s == 96 else 1]
        assert np.array_equal(full.view(np.int64), g[f"hann{s}"].view(np.int64))


def test_patch_grid_order_and_count():
    grid = volume.patch_grid(110, 200, 200, 96)
    assert len(grid) == 18 and grid[0] == (0, 0, 0) and grid[1] == (14, 0, 0) and grid[2] == (0, 0, 52)
    _, origins = ov.make_patches(np.zeros((110, 200, 200), np.float32), 96)
    assert origins == grid
    assert volume.patch_grid(90, 200, 200, 96)[:2] == [(0, 0, 0), (0, 0, 52)]  # D <= P: one z start


def test_patch_indices_cover_everything_once():
    for n, world in [(18, 6), (18, 8), (5, 2), (1, 4)]:
        seen = sorted(i for r in range(world) for i in dist_util.patch_indices(n, r, world))
        assert seen == list(range(n))


def _worker(rank, world, port, n_patches, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    from ddpm3d_b200 import dist_util as du, ensemble, volume as vol
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = du.patch_indices(n_patches)
        local = {i: torch.full((2, 3, 4), float(i)) + torch.arange(4.0) for i in mine}
        got = vol.gather_patches(local, n_patches, device=torch.device("cpu"))
        ok = len(got) == n_patches and all(torch.equal(got[i], torch.full((2, 3, 4), float(i)) + torch.arange(4.0))
                                           for i in range(n_patches))

        class Acc:  # CPU stand-in of ensemble.Welford (whose merge is a CUDA kernel): Chan's pairwise update
            def __init__(self, r):
                g = torch.Generator().manual_seed(100 + r)
                self.x = torch.randn(r + 1, 5, generator=g, dtype=torch.float64)  # rank r holds r + 1 samples
                self.count = r + 1
                self.mean = self.x.mean(0)
                self.m2 = ((self.x - self.mean) ** 2).sum(0)
                self.device = torch.device("cpu")

            def merge(self, mean_b, m2_b, nb):
                n = self.count + nb
                d = mean_b - self.mean
                self.m2 = self.m2 + m2_b + d * d * (self.count * nb / n)
                self.mean = self.mean + d * (nb / n)
                self.count = n

        acc = Acc(rank)
        holds = ensemble.reduce_partials(acc)
        ok = ok and holds == (rank == 0)
        if rank == 0:  # the tree reduction equals the statistics of all samples pooled
            allx = torch.cat([Acc(r).x for r in range(world)])
            ok = ok and acc.count == allx.shape[0]
            ok = ok and torch.allclose(acc.mean, allx.mean(0), atol=1e-12)
            ok = ok and torch.allclose(acc.m2, ((allx - allx.mean(0)) ** 2).sum(0), atol=1e-10)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_patches,world", [(5, 2), (4, 2), (1, 2), (7, 3)])
def test_gloo_gather_and_tree_reduce(n_patches, world):
    """n % world != 0 must not deadlock (the reference's all_gather does); the ensemble's
    Welford partials reduce to rank 0 along a binary tree (world 3: a rank without a partner in round 1)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + n_patches + 10 * world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_patches, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(world))
    assert res == {r: True for r in range(world)}


def test_tiff_and_npz_roundtrip(tmp_path):
    from ddpm3d_b200 import io_formats
    g = np.random.default_rng(1)
    arr = g.random((9, 7, 5), dtype=np.float32)  # (H,W,Z)
    tif = io_formats.write_result(str(tmp_path / "denoised_x.npz"), arr)
    assert np.array_equal(np.load(tmp_path / "denoised_x.npz")["arr_0"], arr)
    back = io_formats.read_volume(tif)  # (Z,H,W)
    assert back.shape == (5, 9, 7) and np.array_equal(back, arr.transpose(2, 0, 1))
    pair = g.random((2, 6, 6, 4), dtype=np.float32)
    np.savez(tmp_path / "pair.npz", pair)
    v = io_formats.read_volume(str(tmp_path / "pair.npz"))
    assert v.shape == (4, 6, 6) and np.allclose(v, (pair[0] / 4).transpose(2, 0, 1))
    with pytest.raises(ValueError):
        io_formats.read_volume("x.png")


def test_slab_bounds():
    from ddpm3d_b200.slab import slab_bounds
    assert slab_bounds(640, 8) == [0, 80, 160, 240, 320, 400, 480, 560, 640]
    assert slab_bounds(110, 8) == [0, 14, 28, 42, 56, 70, 84, 97, 110]
    assert slab_bounds(9, 2) == [0, 5, 9]
    with pytest.raises(ValueError):
        slab_bounds(3, 4)
