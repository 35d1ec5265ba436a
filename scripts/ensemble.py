#!/usr/bin/env python
"""Uncertainty-map ensemble (BASELINE.json config 5; README.md:44 of the reference): N stochastic DDPM samples of one
low-dose patch, seeds rank-strided over the GPUs, voxel-wise mean / variance reduced on the device.

    torchrun --nproc-per-node 8 scripts/ensemble.py --num_samples 16 --timestep_respacing 250 [model / diffusion flags]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch as th  # noqa: E402

from ddpm3d_b200 import dist_util, ensemble, io_formats, volume  # noqa: E402
from ddpm3d_b200.script_util import (add_dict_to_argparser, args_to_dict, sr_create_model_and_diffusion,  # noqa: E402
                                     sr_model_and_diffusion_defaults)


def main():
    defaults = dict(save_dir="", clip_denoised=True, base_samples="", model_path="", num_samples=16, first_seed=10,
                    patch_index=0)
    defaults.update(sr_model_and_diffusion_defaults())
    parser = argparse.ArgumentParser()
    add_dict_to_argparser(parser, defaults)
    args = parser.parse_args()
    dist_util.setup_dist()
    model, diffusion = sr_create_model_and_diffusion(**args_to_dict(args, sr_model_and_diffusion_defaults().keys()))
    if args.model_path:
        model.load_state_dict(dist_util.load_state_dict(args.model_path, map_location="cpu"))
    dev = dist_util.dev()
    model.to(dev)
    if args.use_fp16:
        model.convert_to_fp16()
    model.eval()
    P = args.large_size
    if args.base_samples:
        vol = th.from_numpy(io_formats.read_volume(args.base_samples)).to(dev)
        low = volume.extract_patch(vol, volume.patch_grid(*vol.shape, P)[args.patch_index], P)
    else:  # synthetic low-dose patch
        low = th.rand((1, 1, P, P, P), generator=th.Generator().manual_seed(1234)).to(dev)
    seeds = list(range(args.first_seed, args.first_seed + args.num_samples))
    th.cuda.synchronize()
    t0 = time.time()
    with th.no_grad():
        mean, var, n = ensemble.ensemble_sample(model, diffusion, low, seeds, clip_denoised=args.clip_denoised)
    th.cuda.synchronize()
    dt = time.time() - t0
    if dist_util.get_rank() == 0:
        print(f"ensemble of {n} samples x {diffusion.num_timesteps} steps on {dist_util.get_world_size()} GPU(s): {dt:.1f} s; "
              f"mean in [{float(mean.min()):.3f}, {float(mean.max()):.3f}], mean voxel std {float(var.sqrt().mean()):.4f}", flush=True)
        if args.save_dir:
            os.makedirs(args.save_dir, exist_ok=True)
            np.savez(os.path.join(args.save_dir, "ensemble.npz"), mean=mean[0, 0].cpu().numpy(), variance=var[0, 0].cpu().numpy())
    if th.distributed.is_initialized():
        th.distributed.barrier()
        th.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
