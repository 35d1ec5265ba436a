#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Run in the build container only (it imports
/root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

Every fixture stores OUTPUTS of the reference (guided_diffusion.* imported from
/root/reference) on inputs that `oracle.weights` regenerates from fixed seeds,
so the fixtures stay small.  tests/test_oracle_golden.py pins the oracle to
them; the -m gpu tests pin the CUDA path to them.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from guided_diffusion import gaussian_diffusion as gd  # noqa: E402  (reference)
from guided_diffusion import script_util as su  # noqa: E402  (reference)
from guided_diffusion.nn import timestep_embedding as ref_temb  # noqa: E402
from guided_diffusion.respace import space_timesteps as ref_space  # noqa: E402

from oracle.cases import (  # noqa: E402
    SCHEDULE_CASES, SPACING_CASES, TEMB_CASES, PMV_CASES, UNET_CASES, C1_FLAGS, C1_SHAPE,
    VOLUME_DIMS, VOLUME_Z, HANN_SIZES, UNET2D_CASES,
    sr_flags, cfg_from_flags, model_flags, unet2d_cfg, unet2d_inputs,
)
from oracle.weights import synth_state_dict, synth_inputs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TABLE_NAMES = [
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next",
    "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod",
    "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
    "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2",
]


def ref_model(flags, sd):
    model, diffusion = su.sr_create_model_and_diffusion(**flags)
    model.load_state_dict(sd, strict=True)  # proves key/shape compatibility
    model.eval()
    return model, diffusion


def make_volume_golden():
    """scripts/test.py helpers (the module needs tifffile / mpi4py / blobfile, absent here: stub them; none
    is used by the three functions called)."""
    import importlib.util
    import types
    for name in ("tifffile", "mpi4py", "blobfile"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["mpi4py"].MPI = types.SimpleNamespace()
    spec = importlib.util.spec_from_file_location("ref_test_script", "/root/reference/scripts/test.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    z = {}
    for i, (dim, p, n) in enumerate(VOLUME_DIMS):
        z[f"xy{i}"] = np.array(ref._calculate_xy_starts_fixed(dim, p, num_patches=n), dtype=np.int64)
    for i, (dim, p) in enumerate(VOLUME_Z):
        z[f"z{i}"] = np.array(ref._calculate_z_starts_with_overlap(dim, p), dtype=np.int64)
    for s_ in HANN_SIZES:
        z[f"hann{s_}"] = ref.create_3d_hann_window(s_).reshape(-1)[::7 if s_ == 96 else 1]
    np.savez_compressed(os.path.join(OUT, "volume.npz"), **z)


DDIM_ETAS = [0.0, 0.5, 1.0]


def make_ddim_golden():
    """ddim_sample (gaussian_diffusion.py:537-585) of the reference on the PMV cases' inputs, three etas."""
    z = {}
    for i, case in enumerate(PMV_CASES):
        if case.get("previous_x") or case.get("learned"):
            continue
        d = su.create_gaussian_diffusion(**case["diffusion"])
        g = torch.Generator().manual_seed(100 + i)
        oc = 2 if case["diffusion"].get("learn_sigma") else 1
        shape = (2, 1, 3, 4, 5)
        x = torch.randn(shape, generator=g)
        mo = torch.randn((2, oc, 3, 4, 5), generator=g) * 1.5
        noise = torch.randn(shape, generator=g)
        t = torch.tensor(case["t"])
        for eta in DDIM_ETAS:
            orig = torch.randn_like
            torch.randn_like = lambda _x: noise
            try:
                out = d.ddim_sample(lambda *_a, **_k: mo, x, t, clip_denoised=case["clip"], eta=eta)
            finally:
                torch.randn_like = orig
            z[f"{i}/{eta}/sample"] = out["sample"].numpy()
            z[f"{i}/{eta}/pred_xstart"] = out["pred_xstart"].contiguous().numpy()
    np.savez_compressed(os.path.join(OUT, "ddim.npz"), **z)


def make_unet2d_golden():
    """The model classes scripts/test.py does not instantiate: the 2-D RGB UNetModel from
    create_model_and_diffusion (script_util.py:74-184), the 2-D SuperResModel (unet.py:1654-1673) and a dims=3 UNetModel with its middle-block attention."""
    from guided_diffusion import unet as ref_unet
    z, keys = {}, {}
    for name, case in UNET2D_CASES.items():
        cfg = unet2d_cfg(case)
        sd = synth_state_dict(cfg, seed=case.get("seed", 0))
        if case["kind"] == "create_model":
            model, _ = su.create_model_and_diffusion(**model_flags(**case["flags"]))
        else:
            model = getattr(ref_unet, case["kind"])(**case["ctor"])
        model.load_state_dict(sd, strict=True)
        model.eval()
        keys[name] = [[k, list(v.shape)] for k, v in model.state_dict().items()]
        x, low = unet2d_inputs(case)
        kw = {}
        if low is not None:
            kw["low_res"] = low
        if "y" in case:
            kw["y"] = torch.tensor(case["y"])
        with torch.no_grad():
            z[f"{name}/out"] = model(x, torch.tensor(case["t"]), **kw).numpy()
    np.savez_compressed(os.path.join(OUT, "unet2d.npz"), **z)
    with open(os.path.join(OUT, "state_dict_keys_2d.json"), "w") as f:
        json.dump(keys, f)


def main():
    if "--unet2d-only" in sys.argv:
        make_unet2d_golden()
        return
    if "--volume-only" in sys.argv:
        make_volume_golden()
        return
    if "--ddim-only" in sys.argv:
        make_ddim_golden()
        return
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)

    # ---- schedules / respacing -------------------------------------------------
    z = {}
    for i, kw in enumerate(SCHEDULE_CASES):
        d = su.create_gaussian_diffusion(**kw)
        z[f"{i}/timestep_map"] = np.array(d.timestep_map, dtype=np.int64)
        for n in TABLE_NAMES:
            z[f"{i}/{n}"] = getattr(d, n)
    for i, (n, spec) in enumerate(SPACING_CASES):
        z[f"space{i}"] = np.array(sorted(ref_space(n, spec)), dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "schedules.npz"), **z)

    # ---- timestep embedding ------------------------------------------------------
    z = {}
    for i, (ts, dim) in enumerate(TEMB_CASES):
        z[f"{i}"] = ref_temb(torch.tensor(ts), dim).numpy()
    np.savez_compressed(os.path.join(OUT, "temb.npz"), **z)

    # ---- p_mean_variance / p_sample KATs ---------------------------------------
    z = {}
    for i, case in enumerate(PMV_CASES):
        d = su.create_gaussian_diffusion(**case["diffusion"])
        g = torch.Generator().manual_seed(100 + i)
        C = 1
        oc = 2 * C if case["diffusion"].get("learn_sigma") else C
        shape = (2, C, 3, 4, 5)
        x = torch.randn(shape, generator=g)
        mo = torch.randn((2, oc, 3, 4, 5), generator=g) * 1.5
        noise = torch.randn(shape, generator=g)
        t = torch.tensor(case["t"])
        if case.get("previous_x"):
            d.model_mean_type = gd.ModelMeanType.PREVIOUS_X
        if case.get("learned"):
            d.model_var_type = gd.ModelVarType.LEARNED
        out = d.p_mean_variance(lambda *_a, **_k: mo, x, t, clip_denoised=case["clip"])
        orig = torch.randn_like
        torch.randn_like = lambda _x: noise
        try:
            ps = d.p_sample(lambda *_a, **_k: mo, x, t, clip_denoised=case["clip"])
        finally:
            torch.randn_like = orig
        for k in ("mean", "variance", "log_variance", "pred_xstart"):
            z[f"{i}/{k}"] = out[k].contiguous().numpy()
        z[f"{i}/sample"] = ps["sample"].numpy()
    np.savez_compressed(os.path.join(OUT, "pmv.npz"), **z)

    # ---- UNet forward on tiny configs ------------------------------------------
    z = {}
    keys_meta = {}
    for name, case in UNET_CASES.items():
        flags = sr_flags(**case["flags"])
        cfg = cfg_from_flags(flags)
        sd = synth_state_dict(cfg, seed=case.get("seed", 0))
        model, diffusion = ref_model(flags, sd)
        ref_sd = model.state_dict()
        keys_meta[name] = [[k, list(v.shape)] for k, v in ref_sd.items()]
        low, x, _ = synth_inputs(case["shape"], 0)
        t = torch.tensor(case["t"])
        kw = {}
        if flags["class_cond"]:
            kw["y"] = torch.tensor(case["y"])
        with torch.no_grad():
            out = model(x, t, low_res=low, **kw)
        z[f"{name}/out"] = out.numpy()
    # the two BASELINE configs: key lists only (weights are too big to commit)
    for name, flags in (("C1", sr_flags(**C1_FLAGS)), ("C2", sr_flags())):
        with torch.device("meta"):
            m, _ = su.sr_create_model_and_diffusion(**flags)
        keys_meta[name] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    np.savez_compressed(os.path.join(OUT, "unet_tiny.npz"), **z)
    with open(os.path.join(OUT, "state_dict_keys.json"), "w") as f:
        json.dump(keys_meta, f)

    # ---- C1: full 10-step loop with injected noise ------------------------------
    flags = sr_flags(**C1_FLAGS)
    cfg = cfg_from_flags(flags)
    sd = synth_state_dict(cfg, seed=0)
    model, diffusion = ref_model(flags, sd)
    T = diffusion.num_timesteps
    low, x_T, noises = synth_inputs(C1_SHAPE, T)
    seen_t, mos = [], []
    raw_forward = model.forward

    def spy(x, ts, **kw):
        o = raw_forward(x, ts, **kw)
        seen_t.append(ts.clone())
        mos.append(o.clone())
        return o

    model.forward = spy
    it = iter(noises)
    orig = torch.randn_like
    torch.randn_like = lambda _x: next(it)
    try:
        sample = diffusion.p_sample_loop(model, C1_SHAPE, x_T, clip_denoised=True,
                                         model_kwargs={"low_res": low})
    finally:
        torch.randn_like = orig
    z = {
        "sample": sample.numpy(),
        "model_t": torch.stack(seen_t).numpy(),
        "mo_first": mos[0].numpy(),
        "mo_strided": np.stack([m.numpy().reshape(-1)[::11] for m in mos]),
        "mo_absmax": np.array([float(m.abs().max()) for m in mos]),
        "mo_std": np.array([float(m.std()) for m in mos]),
    }
    np.savez_compressed(os.path.join(OUT, "c1_loop.npz"), **z)
    make_volume_golden()
    make_ddim_golden()
    make_unet2d_golden()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
