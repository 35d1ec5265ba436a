"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, the native topology builder reproduces the reference's state_dict contract, the host
schedule code is bit-exact against the reference fixtures, and the product never touches oracle/."""
import argparse
import ctypes as C
import json
import os
import re

import numpy as np
import pytest
import torch

import ddpm3d_b200 as pkg
from ddpm3d_b200 import _native as N
from ddpm3d_b200 import script_util as su
from ddpm3d_b200.respace import space_timesteps
from oracle import cases
from oracle.schedule import make_tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABLES = [
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next", "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
    "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
    "posterior_mean_coef1", "posterior_mean_coef2",
]


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ddpm3d.h")).read()
    declared = set(re.findall(r"\b(ddpm3d_[a-z0-9_]+)\s*\(", header))
    assert declared, "header parse failed"
    lib = C.CDLL(N.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ddpm3d.h but not exported"
    assert declared == set(N.SIGNATURES), "ctypes binding and header disagree"
    assert N.lib().ddpm3d_abi_version() == 5


def test_struct_layouts_match_header():
    assert C.sizeof(N.StepScalars) == 64
    assert C.sizeof(N.ProfRecord) == 24
    assert C.sizeof(N.Config) == 4 * (6 + 8 + 1 + 8 + 8 + 3)


def build_other_model(case):
    """The model classes of cases.UNET2D_CASES through the drop-in factory / class names."""
    from ddpm3d_b200 import unet
    if case["kind"] == "create_model":
        return su.create_model_and_diffusion(**cases.model_flags(**case["flags"]))[0]
    return getattr(unet, case["kind"])(**case["ctor"])


@pytest.mark.parametrize("name", list(cases.UNET2D_CASES))
def test_state_dict_contract_other_model_classes(golden_dir, name):
    """create_model_and_diffusion's 2-D UNetModel, SuperResModel and a dims=3 UNetModel: keys, order and shapes of
    the native topology == the reference module's state_dict(); forward without a GPU fails loudly."""
    meta = json.load(open(os.path.join(golden_dir, "state_dict_keys_2d.json")))[name]
    case = cases.UNET2D_CASES[name]
    model = build_other_model(case)
    assert [[k, list(v.shape)] for k, v in model.state_dict().items()] == meta
    assert not model._fused_sampler
    if not torch.cuda.is_available():
        x, low = cases.unet2d_inputs(case)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            model(x, torch.tensor(case["t"]), **({"low_res": low} if low is not None else {}))


@pytest.mark.parametrize("name", ["C1", "C2", "tiny", "attn", "plainconv", "classcond", "wide"])
def test_state_dict_contract(golden_dir, name):
    """Keys, order and shapes of the native model == the reference module's state_dict()."""
    meta = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))[name]
    if name == "C1":
        flags = cases.sr_flags(**cases.C1_FLAGS)
    elif name == "C2":
        flags = cases.sr_flags()
    else:
        flags = cases.sr_flags(**cases.UNET_CASES[name]["flags"])
    if name == "C2":
        # 207 M parameters: enumerate through the C ABI only, without materialising weights
        from ddpm3d_b200.unet import _Ctx, _make_config
        cfg = _make_config(image_size=96, model_channels=128, out_channels=2, num_res_blocks=2,
                           attention_resolutions=(0,), channel_mult=(1, 1, 2, 3, 4), num_classes=None, num_heads=4,
                           num_head_channels=64, num_heads_upsample=-1, use_scale_shift_norm=True,
                           resblock_updown=True, use_new_attention_order=False, precision=N.BF16)
        specs = _Ctx(cfg).param_specs()
    else:
        model, _ = su.sr_create_model_and_diffusion(**flags)
        specs = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    assert [[k, list(s)] for k, s in specs] == meta


@pytest.mark.parametrize("i", range(len(cases.SCHEDULE_CASES)))
def test_schedule_tables_bit_exact(golden_dir, i):
    g = np.load(os.path.join(golden_dir, "schedules.npz"))
    d = su.create_gaussian_diffusion(**cases.SCHEDULE_CASES[i])
    assert d.timestep_map == g[f"{i}/timestep_map"].tolist()
    for n in TABLES:
        a, b = getattr(d, n), g[f"{i}/{n}"]
        assert a.dtype == np.float64 and a.shape == b.shape
        assert np.array_equal(a.view(np.int64), b.view(np.int64)), n


@pytest.mark.parametrize("i", range(len(cases.SPACING_CASES)))
def test_space_timesteps_exact(golden_dir, i):
    g = np.load(os.path.join(golden_dir, "schedules.npz"))
    n, spec = cases.SPACING_CASES[i]
    assert sorted(space_timesteps(n, spec)) == g[f"space{i}"].tolist()
    with pytest.raises(ValueError):
        space_timesteps(1000, "ddim999")
    with pytest.raises(ValueError):
        space_timesteps(10, "20")


def test_step_scalars_are_float32_roundings_of_the_tables():
    kw = dict(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="10")
    d = su.create_gaussian_diffusion(**kw)
    tabs = make_tables(**kw)
    s = d.step_scalars()
    assert len(s) == 10
    for i in range(10):
        assert s[i].model_t == float(tabs.timestep_map[i])
        assert np.float32(s[i].sqrt_recip_alphas_cumprod) == np.float32(tabs.sqrt_recip_alphas_cumprod[i])
        assert np.float32(s[i].sqrt_recipm1_alphas_cumprod) == np.float32(tabs.sqrt_recipm1_alphas_cumprod[i])
        assert np.float32(s[i].posterior_mean_coef1) == np.float32(tabs.posterior_mean_coef1[i])
        assert np.float32(s[i].posterior_mean_coef2) == np.float32(tabs.posterior_mean_coef2[i])
        assert np.float32(s[i].min_log) == np.float32(tabs.posterior_log_variance_clipped[i])
        assert np.float32(s[i].max_log) == np.float32(np.log(tabs.betas)[i])
    # rescale_timesteps: respace.py:123-128
    d2 = su.create_gaussian_diffusion(**{**kw, "rescale_timesteps": True, "steps": 4000, "timestep_respacing": "8"})
    m = torch.tensor(d2.timestep_map)
    want = (m.float() * (1000.0 / 4000)).numpy()
    got = np.array([d2.step_scalars()[i].model_t for i in range(8)], dtype=np.float32)
    assert np.array_equal(got, want)


def test_factory_defaults_and_argparse_roundtrip():
    d = su.sr_model_and_diffusion_defaults()
    assert d["large_size"] == 256 and d["small_size"] == 64 and "channel_mult" not in d
    assert "use_new_attention_order" not in d and len(d) == 22
    p = argparse.ArgumentParser()
    su.add_dict_to_argparser(p, dict(clip_denoised=True, batch_size=1, base_samples="", **d))
    a = p.parse_args("--large_size 96 --small_size 96 --num_channels 128 --num_res_blocks 2 --learn_sigma True "
                     "--attention_resolutions 1000 --resblock_updown True --use_fp16 True --num_head_channels 64 "
                     "--use_scale_shift_norm True --diffusion_steps 1000 --noise_schedule linear".split())
    kw = su.args_to_dict(a, d.keys())
    assert kw["use_fp16"] is True and kw["large_size"] == 96 and kw["attention_resolutions"] == "1000"
    with pytest.raises(argparse.ArgumentTypeError):
        su.str2bool("maybe")
    # create_model_and_diffusion (script_util.py:74-127): 23 kwargs, defaults build the 64x64 RGB UNetModel
    md = su.model_and_diffusion_defaults()
    assert len(md) == 23
    m, dif = su.create_model_and_diffusion(**{**md, "num_channels": 32, "num_res_blocks": 1})
    assert m.dims == 2 and m.in_channels == 3 and m.out_channels == 3 and m._middle_attention
    assert dif.num_timesteps == 1000
    with pytest.raises(ValueError):
        su.create_model(48, 32, 1)
    with pytest.raises(NotImplementedError):
        su.create_model(512, 32, 1)


def test_model_surface_without_gpu():
    flags = cases.sr_flags(**cases.UNET_CASES["tiny"]["flags"])
    model, diffusion = su.sr_create_model_and_diffusion(**flags)
    sd = model.state_dict()
    # zero_module parameters start at zero like the reference's (unet.py:210-212,996)
    assert float(sd["out.2.weight"].abs().max()) == 0.0
    assert float(sd["input_blocks.1.0.out_layers.3.weight"].abs().max()) == 0.0
    bad = dict(sd)
    bad.pop("out.2.bias")
    with pytest.raises(RuntimeError):
        model.load_state_dict(bad)
    bad = dict(sd)
    bad["out.2.bias"] = torch.zeros(3)
    with pytest.raises(RuntimeError):
        model.load_state_dict(bad)
    model.load_state_dict(sd)
    assert next(model.parameters()).device.type == "cpu"
    assert model.eval() is model
    x = torch.zeros(1, 1, 4, 16, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(x, torch.tensor([3]), low_res=x)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device error path")
def test_native_calls_fail_loudly_without_a_device():
    from ddpm3d_b200.unet import _Ctx, _make_config
    cfg = _make_config(image_size=16, model_channels=32, out_channels=2, num_res_blocks=1, attention_resolutions=(0,),
                       channel_mult=(1, 1), num_classes=None, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
                       use_scale_shift_norm=True, resblock_updown=True, use_new_attention_order=False, precision=N.FP32)
    ctx = _Ctx(cfg)
    L = N.lib()
    for key, shape in ctx.param_specs():
        t = torch.zeros(shape)
        shp = (C.c_int64 * len(shape))(*shape)
        N.check(L.ddpm3d_load_tensor(ctx, key.encode(), N.ptr(t), shp, len(shape)))
    with pytest.raises(N.NativeError) as e:
        N.check(L.ddpm3d_finalize_weights(ctx, 0))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(N.NativeError):
        N.check(L.ddpm3d_load_tensor(ctx, b"nope.weight", N.ptr(torch.zeros(1)), (C.c_int64 * 1)(1), 1))


def test_bad_config_is_rejected():
    from ddpm3d_b200.unet import _Ctx, _make_config
    kw = dict(image_size=16, model_channels=48, out_channels=2, num_res_blocks=1, attention_resolutions=(0,),
              channel_mult=(1, 1), num_classes=None, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
              use_scale_shift_norm=True, resblock_updown=True, use_new_attention_order=False, precision=N.FP32)
    with pytest.raises(N.NativeError, match="multiple of 32"):
        _Ctx(_make_config(**kw))


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "3d-denoising-diffusion-model_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"
                assert "/root/reference" not in src


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the
    contract's keys, same metric / unit / config family as the GPU arm."""
    import json
    import subprocess
    import sys
    # a small time budget makes the arm fall back from the full 96^3 patch to a z-slab of it (the CPU suite stays short)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, DDPM3D_REF_BUDGET_S="5"))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "unet_evals_per_sec" and line["unit"] == "evals/s"
    from oracle import build_ref
    assert line["cpu_baseline"]["kind"] == ("reference" if build_ref.available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and line["value"] > 0
    assert line["native_so_loaded"] is False  # the reference arm never maps the repo's library
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in line["config"]


def _conv_plan(B, Z, H, W, Cin, Cout, taps=27, extra=0, split_k=1, strip=2, sms=148, dt=None):
    out = (C.c_int32 * 8)()
    N.check(N.lib().ddpm3d_k_conv_plan(N.BF16 if dt is None else dt, B, Z, H, W, Cin, Cout, taps, extra, split_k, strip, sms, out))
    return list(out)


def test_conv_plans_of_the_shipped_network():
    """The planning logic of csrc/conv_tc.cu (plain host arithmetic, no device) on the layer shapes of the shipped
    network (one 96^3 patch, 148 SMs): which tcgen05 kernel runs each resolution class, with which tile.  kind 3 = strip
    kernel {z-planes per tile, positions per tile = MMA N, tiles, grid, weight-ring stages, macro steps, strip rows};
    kind 1 / 2 = brick kernel without / with stream-K {MT, BN, tiles, grid, split-K slots, k-steps}."""
    # 96^2, 48^2 and the 256-channel 24^2 layers: one z-plane per tile, N chosen to tile the padded plane
    assert _conv_plan(1, 96, 96, 96, 128, 128) == [3, 1, 256, 3552, 148, 4, 6, 6]       # 3552 tiles = 24 waves exactly
    assert _conv_plan(1, 96, 96, 96, 128, 128, extra=128)[6] == 8                        # folded skip source: + 2 macro steps
    assert _conv_plan(1, 96, 48, 48, 128, 128) == [3, 1, 240, 960, 148, 4, 6, 8]        # 2400 = 10 x 240 padded positions
    assert _conv_plan(1, 96, 24, 24, 256, 256) == [3, 1, 208, 576, 148, 4, 12, 12]      # 624 = 3 x 208
    # small planes that make one wave: two z-planes per tile share every weight tile
    assert _conv_plan(1, 96, 12, 12, 384, 384) == [3, 2, 176, 144, 144, 4, 18, 16]
    assert _conv_plan(1, 96, 12, 12, 768, 384)[:5] == [3, 2, 176, 144, 144]
    assert _conv_plan(1, 96, 24, 24, 128, 128)[:5] == [3, 2, 208, 144, 144]
    # ... but not when the layer needs several waves, has too few tiles, or the option is at level 1 / 0
    assert _conv_plan(1, 640, 12, 12, 512, 512)[0] == 1
    assert _conv_plan(1, 96, 12, 12, 256, 256)[:4] == [1, 2, 128, 108]
    assert _conv_plan(1, 96, 12, 12, 384, 384, strip=1)[:4] == [2, 2, 128, 162]          # stream-K brick kernel
    assert _conv_plan(1, 96, 12, 12, 384, 384, sms=132)[0] == 2                          # 144 tiles > 132 SMs
    assert _conv_plan(1, 96, 96, 96, 128, 128, strip=0)[:4] == [1, 2, 128, 3456]
    # the 6^2 level: 54 tiles of 128 x 256 for 148 SMs -> k-steps dealt evenly (stream-K), or whole tiles without it
    assert _conv_plan(1, 96, 6, 6, 512, 512)[:5] == [2, 1, 256, 54, 148]
    assert _conv_plan(1, 96, 6, 6, 512, 512, split_k=0)[0] == 1
    # z-slab rank of the 640 x 192 x 192 volume on 8 GPUs: two 96-wide bands per plane
    assert _conv_plan(1, 80, 192, 192, 128, 128)[:5] == [3, 1, 256, 11840, 148]
    # not eligible: channel counts off the 64-element swizzle row, fp32
    assert _conv_plan(1, 96, 12, 12, 100, 384)[0] == 0
    assert _conv_plan(1, 96, 12, 12, 384, 384, dt=N.FP32)[0] == 0
