"""-m gpu: patch extraction, Hann overlap-add and the ensemble reduction kernels against the numpy oracle
(bit-exact: the kernels follow numpy's promotion rules), and the whole volume -> patches -> loop -> blend flow."""
import numpy as np
import pytest
import torch

from ddpm3d_b200 import ensemble, script_util as su, volume
from oracle import cases
from oracle import volume as ov
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu

from gpu_util import DEV  # noqa: E402


@pytest.mark.parametrize("shape,P", [((20, 36, 36), 16), ((16, 16, 16), 16), ((12, 40, 33), 16), ((110, 200, 200), 96)])
def test_extract_and_blend_bit_exact(shape, P):
    g = np.random.default_rng(sum(shape))
    vol = g.random(shape, dtype=np.float32) * 3
    want_patches, origins = ov.make_patches(vol, P)
    assert origins == volume.patch_grid(*shape, P)
    vd = torch.from_numpy(vol).to(DEV)
    got = [volume.extract_patch(vd, o, P)[0, 0] for o in origins]
    for a, b in zip(got, want_patches):
        assert np.array_equal(a.cpu().numpy(), b)
    # blend "denoised" patches (any values) exactly like scripts/test.py:91-139
    den = [torch.from_numpy(g.standard_normal((P, P, P)).astype(np.float32)) for _ in origins]
    want, want_w = ov.blend([d.numpy() for d in den], shape, P)
    arr, wsum = volume.hann_blend(den, origins, shape, P, DEV)
    assert np.array_equal(wsum.cpu().numpy().view(np.int32), want_w.view(np.int32))
    assert np.array_equal(arr.cpu().numpy().view(np.int32), want.view(np.int32))


def test_welford_matches_torch():
    g = torch.Generator().manual_seed(3)
    xs = [torch.randn((1, 1, 6, 8, 8), generator=g) * (1 + k) for k in range(7)]
    acc = ensemble.Welford(xs[0].shape, DEV)
    for x in xs[:4]:
        acc.update(x.to(DEV))
    other = ensemble.Welford(xs[0].shape, DEV)
    for x in xs[4:]:
        other.update(x.to(DEV))
    acc.merge(other.mean, other.m2, other.count)
    st = torch.stack(xs)
    assert acc.count == 7
    assert torch.allclose(acc.mean.cpu(), st.mean(0), rtol=1e-5, atol=1e-6)
    assert torch.allclose(acc.variance().cpu(), st.var(0), rtol=1e-4, atol=1e-5)


def test_denoise_volume_flow_and_ensemble():
    """scripts/test.py main() end to end at a small size: resolution 16, volume (20, 36, 36) -> 18 patches."""
    flags = cases.sr_flags(large_size=16, small_size=16, num_channels=32, num_res_blocks=1, num_head_channels=16,
                           timestep_respacing="3", use_fp16=False)
    cfg = cases.cfg_from_flags(flags)
    model, diffusion = su.sr_create_model_and_diffusion(**flags)
    model.load_state_dict(synth_state_dict(cfg, seed=2))
    model.to(DEV).eval()
    g = np.random.default_rng(0)
    vol = g.random((20, 36, 36), dtype=np.float32)
    out = volume.denoise_volume(model, diffusion, vol, resolution=16)
    assert tuple(out.shape) == (36, 36, 20) and torch.isfinite(out).all()
    # same seed -> same volume; and it equals blending the patches sampled one by one
    out2 = volume.denoise_volume(model, diffusion, vol, resolution=16)
    assert torch.equal(out, out2)
    origins = volume.patch_grid(20, 36, 36, 16)
    torch.cuda.manual_seed_all(10)
    vd = torch.from_numpy(vol).to(DEV)
    pats = []
    for o in origins:
        low = volume.extract_patch(vd, o, 16)
        noise = torch.randn(*low.shape, device=DEV)
        pats.append(diffusion.p_sample_loop(model, tuple(low.shape), noise, model_kwargs={"low_res": low})[0, 0].cpu().numpy())
    want, _ = ov.blend(pats, vol.shape, 16)
    assert np.array_equal(out.cpu().numpy(), want)
    # ensemble: mean / variance over 4 seeds
    low = volume.extract_patch(vd, origins[0], 16)
    mean, var, n = ensemble.ensemble_sample(model, diffusion, low, seeds=[10, 11, 12, 13])
    assert n == 4 and mean.shape == low.shape and float(var.min()) >= 0 and float(var.max()) > 0


def test_ddim_loop_and_script_flags():
    """ddim_sample_loop (gaussian_diffusion.py:625-707) through the native path: eta = 0 is deterministic given x_T
    (no noise consumed matters), equals stepping ddim_sample by hand, and differs from the ancestral sampler."""
    flags = cases.sr_flags(large_size=16, small_size=16, num_channels=32, num_res_blocks=1, num_head_channels=16,
                           timestep_respacing="ddim5", use_fp16=False)
    cfg = cases.cfg_from_flags(flags)
    model, diffusion = su.sr_create_model_and_diffusion(**flags)
    model.load_state_dict(synth_state_dict(cfg, seed=3))
    model.to(DEV).eval()
    shape = (1, 1, 4, 16, 16)
    g = torch.Generator().manual_seed(1)
    x_T = torch.randn(shape, generator=g).to(DEV)
    low = torch.rand(shape, generator=g).to(DEV)
    kw = {"low_res": low}
    a = diffusion.ddim_sample_loop(model, shape, noise=x_T, model_kwargs=kw, eta=0.0)
    b = diffusion.ddim_sample_loop(model, shape, noise=x_T, model_kwargs=kw, eta=0.0)
    assert torch.equal(a, b) and torch.isfinite(a).all()
    img = x_T
    for i in range(diffusion.num_timesteps - 1, -1, -1):
        img = diffusion.ddim_sample(model, img, torch.tensor([i], device=DEV), model_kwargs=kw, eta=0.0)["sample"]
    assert torch.allclose(img, a, rtol=0, atol=1e-6)
    torch.manual_seed(0)
    c = diffusion.p_sample_loop(model, shape, noise=x_T, model_kwargs=kw)
    assert not torch.equal(a, c)
    # the ancestral sampler is unaffected by a previous DDIM call on the same context
    torch.manual_seed(0)
    d = diffusion.p_sample_loop(model, shape, noise=x_T, model_kwargs=kw)
    assert torch.equal(c, d)
