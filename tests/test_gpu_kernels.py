"""-m gpu: every hot-path kernel, called through the C ABI, against the oracle / torch fp32 ops and
the fixtures generated from the unmodified reference."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from ddpm3d_b200 import _native as N
from ddpm3d_b200 import script_util as su
from oracle import cases
from oracle.sampler import p_sample as oracle_p_sample
from oracle.schedule import make_tables

pytestmark = pytest.mark.gpu

from gpu_util import DEV, ROUND_TOL, TDT, conv3d, from_cl, max_rel, pack_weight, stream, to_cl  # noqa: E402


@pytest.mark.parametrize("i", range(len(cases.TEMB_CASES)))
def test_timestep_embedding_matches_reference(golden_dir, i):
    """nn.py:103-121.  With the host frequency table the fp32 angles are bit-identical to the reference's,
    leaving only cosf/sinf: <= 2e-6 absolute.  Without it exp() runs on the device: 1 ulp on the frequency
    times t <= 999 -> <= 1.5e-4 absolute."""
    from ddpm3d_b200.unet import timestep_freqs
    ts, dim = cases.TEMB_CASES[i]
    want = np.load(os.path.join(golden_dir, "temb.npz"))[str(i)]
    t = torch.tensor(ts, dtype=torch.float32, device=DEV)
    out = torch.empty((len(ts), dim), device=DEV)
    fr = timestep_freqs(dim).to(DEV)
    N.check(N.lib().ddpm3d_k_timestep_embedding(N.ptr(t), N.ptr(fr), N.ptr(out), len(ts), dim, stream()))
    torch.cuda.synchronize()
    assert np.abs(out.cpu().numpy() - want).max() <= 2e-6
    N.check(N.lib().ddpm3d_k_timestep_embedding(N.ptr(t), None, N.ptr(out), len(ts), dim, stream()))
    torch.cuda.synchronize()
    assert np.abs(out.cpu().numpy() - want).max() <= 1.5e-4


@pytest.mark.parametrize("dt", [N.FP32, N.BF16, N.FP16])
@pytest.mark.parametrize("C,shape,silu,film,resample", [
    (32, (2, 4, 8, 8), 1, False, 0), (64, (1, 3, 6, 10), 0, True, 0), (128, (1, 8, 16, 16), 1, True, 1),
    (96, (2, 2, 6, 6), 1, False, 2), (384, (1, 4, 12, 12), 1, True, 0), (1024, (1, 2, 6, 6), 1, False, 0),
    (256, (1, 96, 12, 12), 1, True, 0),
])
def test_groupnorm_film_silu(dt, C, shape, silu, film, resample):
    """GroupNorm32 (nn.py:17-19) + FiLM (unet.py:248-252) + SiLU + pool/upsample (unet.py:81-140)."""
    B, Z, H, W = shape
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn((B, C, Z, H, W), generator=g) * 2 + 0.5
    gamma = 1 + 0.1 * torch.randn(C, generator=g)
    beta = 0.1 * torch.randn(C, generator=g)
    fm = torch.randn((B, 2 * C), generator=g) * 0.3 if film else None
    tdt = TDT[dt]
    xin = x.to(tdt).float()  # the kernel sees the rounded input
    ref = F.group_norm(xin, 32, gamma, beta, 1e-5)
    if film:
        ref = ref * (1 + fm[:, :C, None, None, None]) + fm[:, C:, None, None, None]
    if silu:
        ref = F.silu(ref)
    Ho, Wo = H, W
    if resample == 1:
        ref = F.avg_pool3d(ref, (1, 2, 2), (1, 2, 2)); Ho, Wo = H // 2, W // 2
    elif resample == 2:
        ref = F.interpolate(ref, (Z, 2 * H, 2 * W), mode="nearest"); Ho, Wo = 2 * H, 2 * W
    out = torch.empty((B, Z, Ho, Wo, C), device=DEV, dtype=tdt)
    xd, gd_, bd = to_cl(x, tdt), gamma.to(DEV), beta.to(DEV)  # keep the device tensors alive across the call
    fd = fm.to(DEV).contiguous() if film else None
    N.check(N.lib().ddpm3d_k_groupnorm(dt, N.ptr(xd), N.ptr(gd_), N.ptr(bd), N.ptr(fd), silu, resample, N.ptr(out),
                                       B, Z, H, W, C, stream()))
    torch.cuda.synchronize()
    tol = ROUND_TOL[dt]  # one output rounding of the element type
    assert max_rel(from_cl(out), ref) <= tol


@pytest.mark.parametrize("dt", [N.BF16, N.FP16])
@pytest.mark.parametrize("C,shape,silu,film", [(128, (1, 64, 96, 96), 1, True), (256, (1, 47, 81, 81), 1, False),
                                               (384, (2, 30, 48, 48), 0, True)])
def test_groupnorm_large_tensors_streaming_apply(dt, C, shape, silu, film):
    """Tensors of >= 48 MB take the TMA streaming apply kernel (gn_stream.cu): 64-row x 128-channel tiles through a
    shared-memory ring, rewritten in place, stored back by TMA.  Ragged last tile (rows % 64 != 0), batch 2, several
    128-channel chunks, with and without SiLU / FiLM; same bound as the small shapes."""
    B, Z, H, W = shape
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn((B, C, Z, H, W), generator=g) * 2 + 0.5
    gamma = 1 + 0.1 * torch.randn(C, generator=g)
    beta = 0.1 * torch.randn(C, generator=g)
    fm = torch.randn((B, 2 * C), generator=g) * 0.3 if film else None
    tdt = TDT[dt]
    xin = x.to(tdt).float()
    ref = F.group_norm(xin, 32, gamma, beta, 1e-5)
    if film:
        ref = ref * (1 + fm[:, :C, None, None, None]) + fm[:, C:, None, None, None]
    if silu:
        ref = F.silu(ref)
    out = torch.empty((B, Z, H, W, C), device=DEV, dtype=tdt)
    xd, gd_, bd = to_cl(x, tdt), gamma.to(DEV), beta.to(DEV)
    fd = fm.to(DEV).contiguous() if film else None
    N.check(N.lib().ddpm3d_k_groupnorm(dt, N.ptr(xd), N.ptr(gd_), N.ptr(bd), N.ptr(fd), silu, 0, N.ptr(out), B, Z, H, W, C, stream()))
    torch.cuda.synchronize()
    assert max_rel(from_cl(out), ref) <= ROUND_TOL[dt]


@pytest.mark.parametrize("ratio", [100.0, 1000.0])
@pytest.mark.parametrize("C,shape", [(128, (1, 8, 24, 24)), (64, (2, 3, 6, 10)), (384, (1, 96, 12, 12))])
def test_groupnorm_large_mean(ratio, C, shape):
    """|mean| / std = 100 and 1000 per channel (a checkpoint whose activations sit far from zero): E[x^2] - E[x]^2 from
    fp32 sums would cancel (at 1000 the fp32 variance is pure rounding noise).  The statistics kernel sums around a
    per-thread pivot and folds it back in fp64, so the result is as accurate as for zero-mean data: compared with
    F.group_norm evaluated in fp64 on the same fp32 input."""
    B, Z, H, W = shape
    g = torch.Generator().manual_seed(C)
    sign = 1.0 if C != 64 else -1.0
    mean_c = sign * (ratio + 0.2 * torch.randn(C, generator=g))  # every channel of a group sits near +-ratio
    x = torch.randn((B, C, Z, H, W), generator=g) + mean_c[None, :, None, None, None]
    gamma = 1 + 0.1 * torch.randn(C, generator=g)
    beta = 0.1 * torch.randn(C, generator=g)
    ref = F.group_norm(x.double(), 32, gamma.double(), beta.double(), 1e-5)
    out = torch.empty((B, Z, H, W, C), device=DEV)
    xd, gd_, bd = to_cl(x), gamma.to(DEV), beta.to(DEV)
    N.check(N.lib().ddpm3d_k_groupnorm(N.FP32, N.ptr(xd), N.ptr(gd_), N.ptr(bd), None, 0, 0, N.ptr(out), B, Z, H, W, C, stream()))
    torch.cuda.synchronize()
    # the normalised values are O(1); what remains is the fp32 rounding of the per-channel affine y = x A + B at
    # |x A| ~ ratio: a few ratio * 2^-24 (E[x^2] - E[x]^2 from fp32 sums would be off by O(1) at ratio = 1000)
    err = float((from_cl(out).double() - ref).abs().max())
    print(f"groupnorm, |mean|/std = {ratio:.0f}, C = {C}: max abs error {err:.2e}")
    assert err <= 4e-7 * ratio


@pytest.mark.parametrize("dt", [N.BF16, N.FP16])
@pytest.mark.parametrize("bias_scale", [0.0, 100.0])
def test_conv_epilogue_statistics_with_dominant_bias(dt, bias_scale):
    """unet.py:245-247 conv -> GroupNorm32 on the fused path: the tcgen05 conv's epilogue accumulates the channel sums
    of its output and GroupNorm normalises from them.  With a bias of ~100 on every channel of a unit-variance signal
    (|mean| / std = 100) the sums are taken over (x - bias), so nothing cancels: the GroupNorm result equals F.group_norm (fp64) of the
    conv output the kernel itself stored, to one rounding of the 16-bit result."""
    B, Z, H, W, Cin, Cout = 1, 8, 24, 24, 64, 128
    g = torch.Generator().manual_seed(5)
    tdt = TDT[dt]
    x = torch.randn((B, Cin, Z, H, W), generator=g).to(tdt).float()
    w = (torch.randn((Cout, Cin, 3, 3, 3), generator=g) / np.sqrt(Cin * 27)).to(tdt).float()
    b = torch.randn(Cout, generator=g) * 0.1 + bias_scale  # every channel of a group sits near +bias_scale: mean / std ~ 100
    gamma = 1 + 0.1 * torch.randn(Cout, generator=g)
    beta = 0.1 * torch.randn(Cout, generator=g)
    conv_out = torch.empty((B, Z, H, W, Cout), device=DEV, dtype=tdt)
    gn_out = torch.empty_like(conv_out)
    xd, wd, bd, gd_, btd = to_cl(x, tdt), pack_weight(w, tdt), b.to(DEV), gamma.to(DEV), beta.to(DEV)
    N.check(N.lib().ddpm3d_k_conv3d_gn(dt, N.ptr(xd), N.ptr(wd), N.ptr(bd), N.ptr(gd_), N.ptr(btd), N.ptr(conv_out),
                                       N.ptr(gn_out), B, Z, H, W, Cin, Cout, stream()))
    torch.cuda.synchronize()
    assert max_rel(from_cl(conv_out), F.conv3d(x, w, b, padding=1)) <= ROUND_TOL[dt]
    # the statistics describe the fp32 accumulators, the normalisation is applied to their 16-bit rounding: with a
    # dominant bias that rounding (2^-9 / 2^-12 of |x| ~ 150) is what limits the result, not the sums -- so the
    # reference normalises the stored tensor with statistics of the fp32 conv result
    y = F.conv3d(x.double(), w.double(), b.double(), padding=1)
    yg = y.reshape(B, 32, -1)
    mean = yg.mean(-1).repeat_interleave(Cout // 32, 1)[:, :, None, None, None]
    rstd = (yg.var(-1, unbiased=False) + 1e-5).rsqrt().repeat_interleave(Cout // 32, 1)[:, :, None, None, None]
    ref = (from_cl(conv_out).double() - mean) * rstd * gamma.double()[None, :, None, None, None] + beta.double()[None, :, None, None, None]
    err = float((from_cl(gn_out).double() - ref).abs().max() / ref.abs().max())
    print(f"conv -> GroupNorm from epilogue sums, bias scale {bias_scale}: max-rel {err:.2e}")
    assert err <= ROUND_TOL[dt]


@pytest.mark.parametrize("dt", [N.FP32, N.BF16, N.FP16])
@pytest.mark.parametrize("Cin,Cout,shape,taps,stride,res", [
    (2, 32, (1, 4, 8, 8), 27, 1, False), (32, 32, (2, 3, 8, 6), 27, 1, True), (64, 2, (1, 4, 8, 8), 27, 1, False),
    (32, 64, (1, 4, 8, 8), 27, 2, False), (64, 96, (1, 2, 4, 4), 1, 1, True), (128, 128, (1, 5, 12, 12), 27, 1, True),
    (40, 24, (1, 3, 5, 7), 27, 1, False),
])
def test_conv3d_simt(dt, Cin, Cout, shape, taps, stride, res):
    """conv_nd(3, ...) (nn.py:22-32) on the CUDA-core kernel vs F.conv3d fp32."""
    B, Z, H, W = shape
    g = torch.Generator().manual_seed(Cin * 7 + Cout)
    tdt = TDT[dt]
    x = torch.randn((B, Cin, Z, H, W), generator=g).to(tdt).float()
    k = 3 if taps == 27 else 1
    w = (torch.randn((Cout, Cin, k, k, k), generator=g) / np.sqrt(Cin * taps)).to(tdt).float()
    b = torch.randn(Cout, generator=g)
    Ho, Wo = H // stride, W // stride
    r = torch.randn((B, Cout, Z, Ho, Wo), generator=g).to(tdt).float() if res else None
    ref = F.conv3d(x, w, b, stride=(1, stride, stride), padding=k // 2)
    if res:
        ref = ref + r
    out = conv3d(dt, 1, to_cl(x, tdt), pack_weight(w, tdt), b.to(DEV), to_cl(r, tdt) if res else None,
                 B, Z, H, W, Cin, Cout, taps, stride)
    tol = ROUND_TOL[dt]
    assert max_rel(from_cl(out), ref) <= tol


@pytest.mark.parametrize("i", range(len(cases.PMV_CASES)))
def test_p_sample_update_matches_reference(golden_dir, i):
    """gaussian_diffusion.py:232-326,395-439 on the fixtures made by the unmodified reference.
    mean / pred_xstart / log_variance use only non-fused fp32 mul/add: bit-exact.  sample adds exp():
    <= 2 ulp of the noise term."""
    case = cases.PMV_CASES[i]
    g = np.load(os.path.join(golden_dir, "pmv.npz"))
    d = su.create_gaussian_diffusion(**case["diffusion"])
    from ddpm3d_b200 import gaussian_diffusion as gd
    if case.get("previous_x"):
        d.model_mean_type = gd.ModelMeanType.PREVIOUS_X
    if case.get("learned"):
        d.model_var_type = gd.ModelVarType.LEARNED
    gen = torch.Generator().manual_seed(100 + i)
    oc = 2 if case["diffusion"].get("learn_sigma") else 1
    x = torch.randn((2, 1, 3, 4, 5), generator=gen)
    mo = torch.randn((2, oc, 3, 4, 5), generator=gen) * 1.5
    noise = torch.randn((2, 1, 3, 4, 5), generator=gen)
    t = torch.tensor(case["t"], device=DEV)
    out = d._posterior(lambda *a, **k: None, mo.to(DEV), x.to(DEV), t, noise.to(DEV), case["clip"])
    torch.cuda.synchronize()
    for k in ("mean", "pred_xstart", "log_variance"):
        got, want = out[k].cpu().numpy(), g[f"{i}/{k}"]
        assert np.array_equal(got.view(np.int32), want.view(np.int32)), k
    assert np.allclose(out["variance"].cpu().numpy(), g[f"{i}/variance"], rtol=1e-6, atol=0)
    assert np.allclose(out["sample"].cpu().numpy(), g[f"{i}/sample"], rtol=0, atol=1e-6)
    # and the public p_mean_variance with a foreign model callable
    pm = d.p_mean_variance(lambda x_, t_, **k: mo.to(DEV), x.to(DEV), t, clip_denoised=case["clip"])
    assert np.array_equal(pm["mean"].cpu().numpy(), out["mean"].cpu().numpy())


@pytest.mark.parametrize("dt", [N.FP32, N.BF16, N.FP16])
@pytest.mark.parametrize("B,T,C,heads,new_order", [
    (1, 64, 64, 4, 0), (2, 100, 32, 1, 1), (1, 300, 128, 2, 0),
    # 64-wide heads: the fused tcgen05 kernel for 16-bit types (T = 300 exercises the key / query tail masks)
    (2, 256, 64, 1, 1), (1, 1000, 192, 3, 0), (1, 3456, 512, 8, 0),
])
def test_attention_core(dt, B, T, C, heads, new_order):
    """QKVAttentionLegacy / QKVAttention (unet.py:328-393)."""
    g = torch.Generator().manual_seed(T + C)
    tdt = TDT[dt]
    qkv = torch.randn((B, 3 * C, T), generator=g).to(tdt).float()
    ch = C // heads
    s = 1 / np.sqrt(np.sqrt(ch))
    if new_order:
        q, k, v = qkv.chunk(3, dim=1)
        q, k, v = (z.reshape(B * heads, ch, T) for z in (q, k, v))
    else:
        q, k, v = qkv.reshape(B * heads, 3 * ch, T).split(ch, dim=1)
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s), dim=-1)
    ref = torch.einsum("bts,bcs->bct", w, v).reshape(B, C, T)
    out = torch.empty((B, T, C), device=DEV, dtype=tdt)
    qd = qkv.permute(0, 2, 1).contiguous().to(DEV, tdt)
    N.check(N.lib().ddpm3d_k_attention(dt, N.ptr(qd), N.ptr(out),
                                       B, T, C, heads, new_order, stream()))
    torch.cuda.synchronize()
    tol = ROUND_TOL[dt]
    if dt != N.FP32 and C // heads == 64:
        tol *= 3  # tensor-core path: P is rounded to 16 bits before the PV product (like the reference's fp16 einsum)
        simt = torch.empty_like(out)
        N.check(N.lib().ddpm3d_k_attention(dt, N.ptr(qd), N.ptr(simt), B, T, C, heads, new_order | 0x100, stream()))
        torch.cuda.synchronize()
        assert max_rel(out.float().cpu(), simt.float().cpu()) <= tol
    assert max_rel(out.float().permute(0, 2, 1).cpu(), ref) <= tol


@pytest.mark.parametrize("dt", [N.FP32, N.BF16])
@pytest.mark.parametrize("B,T,C,heads,q_begin,q_count", [(1, 300, 64, 4, 100, 150), (2, 512, 128, 2, 256, 256),
                                                          (1, 1000, 192, 3, 872, 128), (2, 384, 64, 1, 0, 192)])
def test_attention_query_window(dt, B, T, C, heads, q_begin, q_count):
    """What a z-slab rank runs after the K/V all-gather: queries [q_begin, q_begin + q_count) against all T keys.
    Rows are independent, so the window must equal the slice of the full result bit for bit (both kernels)."""
    g = torch.Generator().manual_seed(T + q_begin)
    tdt = TDT[dt]
    qd = torch.randn((B, T, 3 * C), generator=g).to(DEV, tdt)
    full = torch.empty((B, T, C), device=DEV, dtype=tdt)
    win = torch.full((B, q_count, C), float("nan"), device=DEV, dtype=tdt)
    N.check(N.lib().ddpm3d_k_attention(dt, N.ptr(qd), N.ptr(full), B, T, C, heads, 0, stream()))
    N.check(N.lib().ddpm3d_k_attention_window(dt, N.ptr(qd), N.ptr(win), B, T, C, heads, 0, q_begin, q_count, stream()))
    torch.cuda.synchronize()
    assert torch.equal(win, full[:, q_begin:q_begin + q_count])
    with pytest.raises(RuntimeError):
        N.check(N.lib().ddpm3d_k_attention_window(dt, N.ptr(qd), N.ptr(win), B, T, C, heads, 0, T - 1, q_count, stream()))


@pytest.mark.parametrize("i", [i for i, c in enumerate(cases.PMV_CASES) if not (c.get("previous_x") or c.get("learned"))])
def test_ddim_update_bit_exact(golden_dir, i):
    """ddim_sample (gaussian_diffusion.py:537-585) against the unmodified reference's outputs.  pred_xstart is
    bit-identical.  The sample goes through three fp32 sqrt(): the kernel's (like torch's CUDA sqrt) is correctly
    rounded, but the fixtures were produced by torch's CPU sqrt, which is 1 ulp off for some arguments (e.g.
    sqrt(4.1181935e-05f)), so the sample is compared to 2 ulp."""
    case = cases.PMV_CASES[i]
    g = np.load(os.path.join(golden_dir, "ddim.npz"))
    d = su.create_gaussian_diffusion(**case["diffusion"])
    gen = torch.Generator().manual_seed(100 + i)
    oc = 2 if case["diffusion"].get("learn_sigma") else 1
    x = torch.randn((2, 1, 3, 4, 5), generator=gen)
    mo = torch.randn((2, oc, 3, 4, 5), generator=gen) * 1.5
    noise = torch.randn((2, 1, 3, 4, 5), generator=gen)
    t = torch.tensor(case["t"], device=DEV)
    for eta in (0.0, 0.5, 1.0):
        out = d.ddim_sample(lambda x_, t_, **k: mo.to(DEV), x.to(DEV), t, clip_denoised=case["clip"], eta=eta,
                            noise=noise.to(DEV))
        torch.cuda.synchronize()
        assert np.allclose(out["sample"].cpu().numpy(), g[f"{i}/{eta}/sample"], rtol=2.4e-7, atol=1e-7), eta
        assert np.array_equal(out["pred_xstart"].cpu().numpy().view(np.int32), g[f"{i}/{eta}/pred_xstart"].view(np.int32))
