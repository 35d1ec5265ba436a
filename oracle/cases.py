"""Oracle: the named test cases shared by make_golden.py and tests/.

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
from __future__ import annotations

from .unet import UNetConfig

# create_gaussian_diffusion kwargs (script_util.py:578-616)
SCHEDULE_CASES = [
    dict(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing=""),
    dict(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="100"),
    dict(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="10"),
    dict(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="250"),
    dict(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="ddim25"),
    dict(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="10,10,10"),
    dict(steps=1000, learn_sigma=False, noise_schedule="cosine", timestep_respacing="50"),
    dict(steps=4000, learn_sigma=True, noise_schedule="cosine", timestep_respacing=""),
    dict(steps=100, learn_sigma=False, sigma_small=True, noise_schedule="linear", timestep_respacing="7"),
    dict(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="1000"),
]

SPACING_CASES = [
    (1000, "250"), (1000, "ddim25"), (1000, "10,10,10"), (300, [10, 15, 20]), (1000, "1"),
    (1000, "999"), (7, "3,2"), (1000, "ddim50"), (1000, [1000]), (50, "50"),
]

TEMB_CASES = [
    ([0, 1, 37, 999], 128),
    ([5, 888], 32),
    ([3, 14], 33),
    ([0, 111, 222, 333, 444, 555, 666, 777, 888, 999], 128),
]

_D = dict(steps=1000, noise_schedule="linear", timestep_respacing="10")
PMV_CASES = [
    dict(diffusion=dict(learn_sigma=True, **_D), t=[5, 5], clip=True),
    dict(diffusion=dict(learn_sigma=True, **_D), t=[0, 9], clip=True),
    dict(diffusion=dict(learn_sigma=True, **_D), t=[3, 0], clip=False),
    dict(diffusion=dict(learn_sigma=False, **_D), t=[0, 7], clip=True),
    dict(diffusion=dict(learn_sigma=False, sigma_small=True, **_D), t=[1, 0], clip=True),
    dict(diffusion=dict(learn_sigma=True, predict_xstart=True, **_D), t=[2, 8], clip=True),
    dict(diffusion=dict(learn_sigma=True, **_D), t=[4, 0], clip=True, previous_x=True),
    dict(diffusion=dict(learn_sigma=True, **_D), t=[6, 1], clip=False, learned=True),
    dict(diffusion=dict(learn_sigma=True, steps=1000, noise_schedule="linear", timestep_respacing=""),
         t=[999, 500], clip=True),
]


def sr_flags(**over) -> dict:
    """test_DDPM_3d_mpi.sh:2-3 launcher flags on top of
    sr_model_and_diffusion_defaults (script_util.py:269-277) -> the 22 kwargs of
    sr_create_model_and_diffusion (script_util.py:280-303)."""
    f = dict(
        large_size=96, small_size=96, class_cond=False, learn_sigma=True, num_channels=128,
        num_res_blocks=2, num_heads=4, num_head_channels=64, num_heads_upsample=-1,
        attention_resolutions="1000", dropout=0.0, diffusion_steps=1000, noise_schedule="linear",
        timestep_respacing="", use_kl=False, predict_xstart=False, rescale_timesteps=False,
        rescale_learned_sigmas=False, use_checkpoint=False, use_scale_shift_norm=True,
        resblock_updown=True, use_fp16=False,
    )
    f.update(over)
    return f


def cfg_from_flags(flags: dict) -> UNetConfig:
    return UNetConfig.from_sr_flags(**flags)


# BASELINE.json configs[0]
C1_FLAGS = dict(large_size=32, small_size=32, num_channels=32, num_res_blocks=1, timestep_respacing="10")
C1_SHAPE = (1, 1, 32, 32, 32)

_TINY = dict(large_size=16, small_size=16, num_channels=32, num_res_blocks=1, num_head_channels=16,
             timestep_respacing="10")
UNET_CASES = {
    # shipped topology, shrunk
    "tiny": dict(flags=dict(**_TINY), shape=(1, 1, 8, 16, 16), t=[888]),
    # batch 2, different t per sample, non-cubic
    "tiny_b2": dict(flags=dict(**_TINY), shape=(2, 1, 4, 16, 32), t=[111, 999], seed=3),
    # attention at ds=2 and 4 (legacy qkv order, unet.py:328-358)
    "attn": dict(flags=dict(**{**_TINY, "attention_resolutions": "8,4"}), shape=(1, 1, 4, 16, 16), t=[444]),
    # factory defaults for the flags the launcher overrides: strided-conv down / conv up, additive emb
    "plainconv": dict(flags=dict(**{**_TINY, "resblock_updown": False, "use_scale_shift_norm": False,
                                    "learn_sigma": False}), shape=(1, 1, 4, 16, 16), t=[0], seed=5),
    # class conditional
    "classcond": dict(flags=dict(**{**_TINY, "class_cond": True}), shape=(2, 1, 4, 16, 16), t=[222, 777],
                      y=[3, 998], seed=7),
    # two res blocks + 64-wide (channel counts 64..256 like the tcgen05 path needs)
    "wide": dict(flags=dict(**{**_TINY, "num_channels": 64, "num_res_blocks": 2, "num_head_channels": 64}),
                 shape=(1, 1, 8, 16, 16), t=[555], seed=9),
}


def model_flags(**over) -> dict:
    """model_and_diffusion_defaults (script_util.py:43-65) -> the 23 kwargs of create_model_and_diffusion
    (script_util.py:74-98), which builds the 2-D RGB `UNetModel` (middle-block attention)."""
    f = dict(
        image_size=64, class_cond=False, learn_sigma=False, num_channels=128, num_res_blocks=2, channel_mult="",
        num_heads=4, num_head_channels=-1, num_heads_upsample=-1, attention_resolutions="16,8", dropout=0.0,
        diffusion_steps=1000, noise_schedule="linear", timestep_respacing="", use_kl=False, predict_xstart=False,
        rescale_timesteps=False, rescale_learned_sigmas=False, use_checkpoint=False, use_scale_shift_norm=True,
        resblock_updown=False, use_fp16=False, use_new_attention_order=False,
    )
    f.update(over)
    return f


# The other model classes of unet.py.  `kind`:
#   "create_model": script_util.create_model_and_diffusion(**model_flags(**flags))  -> 2-D UNetModel
#   "UNetModel" / "SuperResModel": the class instantiated directly with `ctor` (dims 2 or 3)
UNET2D_CASES = {
    "rgb": dict(kind="create_model", flags=dict(num_channels=32, num_res_blocks=1), shape=(2, 3, 64, 64), t=[10, 900]),
    "rgb_opts": dict(kind="create_model",
                     flags=dict(image_size=32, channel_mult="1,2", num_channels=64, num_res_blocks=1, learn_sigma=True,
                                class_cond=True, num_head_channels=32, resblock_updown=True,
                                use_new_attention_order=True, attention_resolutions="16"),
                     shape=(1, 3, 32, 32), t=[333], y=[17], seed=2),
    "sr2d": dict(kind="SuperResModel",
                 ctor=dict(image_size=32, in_channels=3, model_channels=32, out_channels=3, num_res_blocks=1,
                           attention_resolutions=(2,), channel_mult=(1, 2), num_heads=2, use_scale_shift_norm=True),
                 shape=(1, 3, 32, 32), low_shape=(1, 3, 32, 32), t=[600], seed=3),
    "mid3d": dict(kind="UNetModel",
                  ctor=dict(image_size=16, in_channels=2, model_channels=32, out_channels=2, num_res_blocks=1,
                            attention_resolutions=(1000,), channel_mult=(1, 2), dims=3, num_head_channels=16,
                            use_scale_shift_norm=True, resblock_updown=True),
                  shape=(1, 2, 4, 16, 16), t=[50], seed=4),
}


def unet2d_cfg(case) -> UNetConfig:
    if case["kind"] == "create_model":
        return UNetConfig.from_model_flags(**model_flags(**case["flags"]))
    c = dict(case["ctor"])
    sr = case["kind"] == "SuperResModel"
    return UNetConfig(
        image_size=c["image_size"], in_channels=c["in_channels"], model_channels=c["model_channels"],
        out_channels=c["out_channels"], num_res_blocks=c["num_res_blocks"],
        attention_ds=tuple(c["attention_resolutions"]), channel_mult=tuple(c["channel_mult"]),
        num_classes=c.get("num_classes"), num_heads=c.get("num_heads", 1),
        num_head_channels=c.get("num_head_channels", -1), num_heads_upsample=c.get("num_heads_upsample", -1),
        use_scale_shift_norm=c.get("use_scale_shift_norm", False), resblock_updown=c.get("resblock_updown", False),
        use_new_attention_order=c.get("use_new_attention_order", False), dims=c.get("dims", 2),
        middle_attention=True, concat_low_res=sr)


def unet2d_inputs(case):
    """x ~ N(0,1), low_res ~ U[0,1) (SuperResModel only; unet.py:1666-1673 concatenates it as is -- the bilinear
    up-sampling of the upstream code is commented out there), from a CPU generator."""
    import torch
    g = torch.Generator().manual_seed(4321 + case.get("seed", 0))
    x = torch.randn(case["shape"], generator=g)
    low = torch.rand(case["low_shape"], generator=g) if "low_shape" in case else None
    return x, low


# scripts/test.py tiling helpers: (dim, patch, num_patches), (dim, patch), window sizes
VOLUME_DIMS = [(200, 96, 3), (36, 16, 3), (16, 16, 3), (50, 16, 3), (130, 96, 3), (20, 16, 1)]
VOLUME_Z = [(110, 96), (96, 96), (20, 16), (90, 96), (130, 96), (16, 16)]
HANN_SIZES = [96, 16, 5]
