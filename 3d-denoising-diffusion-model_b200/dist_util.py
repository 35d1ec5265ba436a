"""One process per GPU (torchrun) instead of the reference's MPI rendez-vous
(guided_diffusion/dist_util.py:22-78).  Same function names."""
from __future__ import annotations

import os


def setup_dist(backend=None):
    """dist_util.py:22-47.  Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* from the environment
    (torchrun); a single process needs no process group."""
    import torch
    import torch.distributed as dist
    if dist.is_initialized() or int(os.environ.get("WORLD_SIZE", "1")) == 1:
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend=backend, init_method="env://")


def dev():
    """dist_util.py:50-56 (without the hard-coded GPUS_PER_NODE = 2)."""
    import torch
    if torch.cuda.is_available():
        return torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    return torch.device("cpu")


def get_rank():
    return int(os.environ.get("RANK", "0"))


def get_world_size():
    return int(os.environ.get("WORLD_SIZE", "1"))


def load_state_dict(path, **kwargs):
    """dist_util.py:58-78: every rank reads the checkpoint itself (no MPI byte broadcast)."""
    import torch
    return torch.load(path, **kwargs)


def patch_indices(n_patches, rank=None, world=None):
    """scripts/test.py:235-246: patches are rank-strided.  Unlike the reference (which deadlocks in
    all_gather when n_patches % world != 0) the caller gets its own list and gathers by index."""
    rank = get_rank() if rank is None else rank
    world = get_world_size() if world is None else world
    return list(range(rank, n_patches, world))
