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


# scripts/test.py tiling helpers: (dim, patch, num_patches), (dim, patch), window sizes
VOLUME_DIMS = [(200, 96, 3), (36, 16, 3), (16, 16, 3), (50, 16, 3), (130, 96, 3), (20, 16, 1)]
VOLUME_Z = [(110, 96), (96, 96), (20, 16), (90, 96), (130, 96), (16, 16)]
HANN_SIZES = [96, 16, 5]
