#!/usr/bin/env python
"""B200 counterpart of the reference's scripts/test.py: low-dose volume -> 96^3 patches -> full DDPM
ancestral sampling of every patch (rank-strided over GPUs) -> Hann overlap-add -> denoised .npz / .tif.

    torchrun --nproc-per-node 8 scripts/test.py $MODEL_FLAGS $DIFFUSION_FLAGS $SAMPLE_FLAGS

accepts the flag set of test_DDPM_3d_mpi.sh:2-5 (including --num_samples, which the reference's current
test.py rejects) and, besides .tif/.tiff, the README's .npz input.  `--model_path ""` runs with the
constructor's initial weights (for smoke runs without a checkpoint)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch as th  # noqa: E402

from ddpm3d_b200 import dist_util, io_formats, volume  # noqa: E402
from ddpm3d_b200.script_util import (add_dict_to_argparser, args_to_dict, sr_create_model_and_diffusion,  # noqa: E402
                                     sr_model_and_diffusion_defaults)


def create_argparser():
    defaults = dict(save_dir="", clip_denoised=True, batch_size=1, use_ddim=False, eta=0.0, timestep_respacing="",
                    base_samples="", model_path="", num_samples=1, rng="torch")
    defaults.update(sr_model_and_diffusion_defaults())
    parser = argparse.ArgumentParser()
    add_dict_to_argparser(parser, defaults)
    return parser


def log(msg):
    if dist_util.get_rank() == 0:
        print(msg, flush=True)


def main():
    args = create_argparser().parse_args()
    dist_util.setup_dist()
    log("creating model...")
    model, diffusion = sr_create_model_and_diffusion(**args_to_dict(args, sr_model_and_diffusion_defaults().keys()))
    if args.model_path:
        model.load_state_dict(dist_util.load_state_dict(args.model_path, map_location="cpu"))
    dev = dist_util.dev()
    model.to(dev)
    if args.use_fp16:
        model.convert_to_fp16()
    model.eval()
    log("loading data...")
    vol = io_formats.read_volume(args.base_samples)
    log(f"Using original data without normalization - min: {vol.min():.4f}, max: {vol.max():.4f}, std: {vol.std():.4f}")
    t0 = time.time()
    sample_fn = None
    if args.use_ddim:  # scripts/test_backup.py:63-71: same network, DDIM update with eta
        def sample_fn(low_res):
            shape = tuple(low_res.shape)
            noise = th.randn(*shape, device=low_res.device)
            return diffusion.ddim_sample_loop(model, shape, noise, clip_denoised=args.clip_denoised,
                                              model_kwargs={"low_res": low_res}, eta=args.eta, rng=args.rng)
    with th.no_grad():
        arr = volume.denoise_volume(model, diffusion, vol, resolution=args.large_size, clip_denoised=args.clip_denoised,
                                    seed=10, log=log, sample_fn=sample_fn, **({} if args.use_ddim else {"rng": args.rng}))
    th.cuda.synchronize()
    log(f"sampling + blending took {time.time() - t0:.1f} s")
    if dist_util.get_rank() == 0:
        out_dir = args.save_dir or os.getcwd()
        os.makedirs(out_dir, exist_ok=True)
        base = os.path.basename(args.base_samples)
        for ext in (".tiff", ".tif", ".npz", ".npy"):
            base = base.replace(ext, "")
        out_path = os.path.join(out_dir, f"denoised_{base}.npz")
        tif = io_formats.write_result(out_path, arr.cpu().numpy())
        log(f"saved {out_path} and {tif}")
    if th.distributed.is_initialized():
        th.distributed.barrier()
        th.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
