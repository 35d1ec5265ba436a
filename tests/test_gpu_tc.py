"""-m gpu: the tcgen05 implicit-GEMM convolution (csrc/conv_tc.cu) against F.conv3d and against the
CUDA-core kernel on identical bf16 inputs, plus whole-network parity with the tensor-core path forced."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from ddpm3d_b200 import _native as N
from oracle import cases
from oracle.weights import synth_inputs

pytestmark = pytest.mark.gpu

from gpu_util import DEV, ROUND_TOL, TDT, conv3d, from_cl, max_rel, pack_weight, to_cl  # noqa: E402
from test_gpu_model import TOL, build  # noqa: E402


@pytest.mark.parametrize("Cin,Cout,shape,taps,res", [
    (64, 64, (1, 4, 8, 8), 27, False),       # one 128-row brick per z-pair
    (64, 128, (1, 2, 16, 16), 27, True),     # BN=128
    (128, 128, (2, 5, 12, 12), 27, True),    # batch 2, bricks overhang the volume (12 not a power of 2)
    (128, 64, (1, 3, 6, 6), 27, False),      # 6x6 plane: 108-row bricks
    (192, 384, (1, 4, 12, 12), 27, True),    # 3 K chunks per tap, 3 N tiles
    (256, 128, (1, 8, 24, 24), 27, False),   # more tiles than one wave of k-steps
    (64, 192, (1, 4, 8, 8), 1, True),        # 1x1x1 (attention qkv / proj)
    (128, 128, (1, 96, 6, 6), 27, True),     # the shipped lowest resolution (Z = 96)
    (128, 128, (1, 7, 48, 48), 27, False),   # odd Z
    (128, 128, (1, 24, 96, 96), 27, True),   # two bricks per CTA sharing the weight tile (MT = 2), paired along h
    (256, 256, (1, 48, 24, 24), 27, True),   # BN = 256
    (256, 128, (1, 21, 48, 48), 27, False),  # MT = 2 with an odd brick count (edge brick fully out of bounds)
    (512, 512, (1, 96, 6, 6), 27, False),    # small-M layer: split-K (fp32 partials + reduce pass)
    (384, 384, (1, 96, 12, 12), 27, True),   # split-K with a residual applied in the reduce pass
    (768, 384, (1, 24, 12, 12), 27, False),  # long K (324 k-steps), few tiles
    # strip variant (A staged once per (dz, chunk), 9 taps through row-shifted descriptors): >= 296 tiles
    (128, 128, (2, 16, 96, 96), 27, True),   # batch 2, residual
    (64, 128, (1, 40, 48, 48), 27, False),   # 48-wide band (pitch 50)
    (128, 256, (1, 32, 48, 48), 27, True),   # two output-channel tiles
    (64, 128, (1, 4, 192, 192), 27, False),  # two 96-wide bands per plane (halo column shared between bands)
    (64, 128, (1, 12, 80, 80), 27, True),    # band width that is not a power of two
    # strip variant on small planes ("strip" level 2): tiles of 160 / 176 padded-flattened positions, two z-planes per
    # tile sharing every weight tile (their accumulators fill the 512 TMEM columns)
    (128, 384, (1, 48, 16, 16), 27, False),  # 16-wide planes: two 160-position tiles per plane, the second one ragged
    (128, 128, (1, 145, 20, 12), 27, True),  # 12-wide, 20 rows: N = 160, residual, odd Z (the last pair has one plane)
    (192, 384, (1, 95, 12, 12), 27, True),   # 12 x 12 planes, N = 176, three channel tiles (the shipped 12^2 level), odd Z
    (128, 384, (2, 48, 12, 12), 27, False),  # batch 2
])
@pytest.mark.parametrize("dt", [N.BF16, N.FP16])
def test_conv3d_tcgen05(Cin, Cout, shape, taps, res, dt):
    B, Z, H, W = shape
    tdt = TDT[dt]
    g = torch.Generator().manual_seed(Cin + 3 * Cout + H)
    x = torch.randn((B, Cin, Z, H, W), generator=g).to(tdt).float()
    k = 3 if taps == 27 else 1
    w = (torch.randn((Cout, Cin, k, k, k), generator=g) / np.sqrt(Cin * taps)).to(tdt).float()
    b = torch.randn(Cout, generator=g)
    r = torch.randn((B, Cout, Z, H, W), generator=g).to(tdt).float() if res else None
    ref = F.conv3d(x, w, b, padding=k // 2)
    if res:
        ref = ref + r
    args = (to_cl(x, tdt), pack_weight(w, tdt), b.to(DEV), to_cl(r, tdt) if res else None, B, Z, H, W, Cin, Cout, taps, 1)
    tc = conv3d(dt, 2, *args)
    simt = conv3d(dt, 1, *args)
    assert max_rel(from_cl(tc), ref) <= ROUND_TOL[dt]
    # the same layer with the strip variant off (brick / stream-K kernel) and restricted to the large layers
    for path in (5, 6):
        assert max_rel(conv3d(dt, path, *args).float().cpu(), simt.float().cpu()) <= 1.5 * ROUND_TOL[dt]
    # same 16-bit inputs, fp32 accumulation in both: they may differ by one output rounding at most
    assert max_rel(tc.float().cpu(), simt.float().cpu()) <= 1.5 * ROUND_TOL[dt]
    frac_equal = float((tc == simt).float().mean())
    # fp16 keeps 3 more mantissa bits, so fp32 summation-order differences flip the last bit more often
    assert frac_equal >= (0.98 if dt == N.BF16 else 0.90), frac_equal


@pytest.mark.parametrize("Cin,Cout,shape", [(64, 64, (1, 4, 16, 16)), (128, 128, (2, 5, 24, 24)), (256, 256, (1, 9, 12, 20)),
                                            (128, 128, (1, 24, 96, 96)), (64, 128, (1, 3, 10, 6))])
@pytest.mark.parametrize("dt", [N.BF16, N.FP16])
def test_strided_downsample_conv_tcgen05(Cin, Cout, shape, dt):
    """Downsample(use_conv=True) = Conv3d(C, C, 3, stride=(1,2,2), padding=1) (unet.py:113-140; the factory default
    resblock_updown=False) on the tcgen05 kernel: the A operand is a TMA box with element strides (2, 2) on (W, H)."""
    B, Z, H, W = shape
    tdt = TDT[dt]
    g = torch.Generator().manual_seed(Cin + Cout + H)
    x = torch.randn((B, Cin, Z, H, W), generator=g).to(tdt).float()
    w = (torch.randn((Cout, Cin, 3, 3, 3), generator=g) / np.sqrt(Cin * 27)).to(tdt).float()
    b = torch.randn(Cout, generator=g)
    ref = F.conv3d(x, w, b, stride=(1, 2, 2), padding=1)
    args = (to_cl(x, tdt), pack_weight(w, tdt), b.to(DEV), None, B, Z, H, W, Cin, Cout, 27, 2)
    tc = conv3d(dt, 2, *args)
    simt = conv3d(dt, 1, *args)
    assert max_rel(from_cl(tc), ref) <= ROUND_TOL[dt]
    assert max_rel(tc.float().cpu(), simt.float().cpu()) <= 1.5 * ROUND_TOL[dt]


@pytest.mark.parametrize("Cin,Cout,shape,stride", [(64, 64, (2, 1, 32, 32), 1), (128, 256, (1, 1, 64, 64), 1), (64, 128, (1, 3, 16, 24), 1),
                                                   (128, 128, (1, 1, 32, 32), 2), (64, 128, (1, 4, 160, 192), 1)])
@pytest.mark.parametrize("dt", [N.BF16, N.FP16])
def test_conv2d_nine_taps(Cin, Cout, shape, stride, dt):
    """The Conv2d layers of the dims = 2 networks (unet.py:396-716): 9 in-plane taps on one-plane volumes (Z > 1 =
    the same 3x3 kernel on every plane), on the tcgen05 kernels (brick, strip for the last shape, element-strided
    boxes for stride 2) and on the CUDA-core kernel, against F.conv2d."""
    B, Z, H, W = shape
    tdt = TDT[dt]
    g = torch.Generator().manual_seed(Cin + Cout + W)
    x = torch.randn((B, Cin, Z, H, W), generator=g).to(tdt).float()
    w = (torch.randn((Cout, Cin, 3, 3), generator=g) / np.sqrt(Cin * 9)).to(tdt).float()
    b = torch.randn(Cout, generator=g)
    ref = torch.stack([F.conv2d(x[:, :, z], w, b, stride=stride, padding=1) for z in range(Z)], dim=2)
    args = (to_cl(x, tdt), pack_weight(w, tdt), b.to(DEV), None, B, Z, H, W, Cin, Cout, 9, stride)
    tc = conv3d(dt, 2, *args)
    simt = conv3d(dt, 1, *args)
    assert max_rel(from_cl(tc), ref) <= ROUND_TOL[dt]
    assert max_rel(from_cl(simt), ref) <= ROUND_TOL[dt]


def test_ineligible_shapes_are_rejected():
    x = torch.zeros((1, 2, 4, 4, 32), device=DEV, dtype=torch.bfloat16)
    w = torch.zeros((64, 27 * 32), device=DEV, dtype=torch.bfloat16)
    b = torch.zeros(64, device=DEV)
    with pytest.raises(N.NativeError, match="not eligible"):
        conv3d(N.BF16, 2, x, w, b, None, 1, 2, 4, 4, 32, 64)


@pytest.mark.parametrize("half", [True, "fp16"])
@pytest.mark.parametrize("name", ["wide"])
def test_unet_on_tensor_cores_matches_simt(golden_dir, name, half):
    """Every eligible convolution on tcgen05 (skip folding, pooled / upsampled residuals,
    channel-concat sources) against the same network on the CUDA-core kernels."""
    case = cases.UNET_CASES[name]
    low, x, _ = synth_inputs(case["shape"], 0)
    outs = {}
    for path in (1, 2):
        model, _, _, _ = build(case["flags"], seed=case.get("seed", 0), fp16=half)
        model.set_option("conv_path", path)
        model.set_option("profile", 1)
        outs[path] = model(x.to(DEV), torch.tensor(case["t"], device=DEV), low_res=low.to(DEV)).cpu()
        kinds = {k for k, _, _ in model.profile_read()}
        assert ("conv_tcgen05" in kinds) == (path == 2)
    # two equally valid bf16 evaluations diverge by the bf16 rounding floor of the network (see test_gpu_model.TOL)
    assert max_rel(outs[2], outs[1]) <= TOL[half]


def test_fused_groupnorm_statistics_agree_with_the_separate_pass():
    """GroupNorm statistics taken from the conv epilogue's fp32 accumulators (per-CTA channel sums) vs the
    separate statistics pass over the stored 16-bit tensor: same network output up to the rounding floor, and
    bit-identical from run to run (fixed reduction order)."""
    case = cases.UNET_CASES["wide"]
    low, x, _ = synth_inputs(case["shape"], 0)
    outs = {}
    for fuse in (1, 0):
        model, _, _, _ = build(case["flags"], seed=case.get("seed", 0), fp16="fp16")
        model.set_option("fuse_stats", fuse)
        a = model(x.to(DEV), torch.tensor(case["t"], device=DEV), low_res=low.to(DEV)).cpu()
        b = model(x.to(DEV), torch.tensor(case["t"], device=DEV), low_res=low.to(DEV)).cpu()
        assert torch.equal(a, b)
        outs[fuse] = a
    assert max_rel(outs[1], outs[0]) <= 3e-3


def test_fused_statistics_on_the_shipped_architecture():
    """The same on the shipped network with 96 x 96 planes (8 of them): here the stem's tensor-core kernel, the strip
    kernel with a nearest-upsampled residual (up ResBlocks at the top level) and the brick kernels all hand channel sums
    to the GroupNorm that follows; only the stream-K layers still take the statistics pass."""
    from oracle.weights import synth_state_dict
    from ddpm3d_b200 import script_util as su
    flags = cases.sr_flags(use_fp16=True)
    sd = synth_state_dict(cases.cfg_from_flags(flags), seed=4)
    shape = (1, 1, 8, 96, 96)
    low, x, _ = synth_inputs(shape, 0)
    t = torch.tensor([321], device=DEV)
    outs, n_launch = {}, {}
    for fuse in (1, 0):
        model, _ = su.sr_create_model_and_diffusion(**flags)
        model.load_state_dict(sd)
        model.to(DEV)
        model.set_half_dtype("fp16")
        model.convert_to_fp16()
        model.eval()
        model.set_option("fuse_stats", fuse)
        a = model(x.to(DEV), t, low_res=low.to(DEV)).cpu()
        b = model(x.to(DEV), t, low_res=low.to(DEV)).cpu()
        assert torch.equal(a, b)
        outs[fuse] = a
    err = max_rel(outs[1], outs[0])
    print(f"C2 architecture, fused vs separate GroupNorm statistics: max-rel {err:.2e}")
    assert err <= 3e-3


@pytest.mark.parametrize("half", [True, "fp16", "bf16_strict"])
@pytest.mark.parametrize("shape,learn_sigma,ch", [((1, 1, 7, 16, 32), True, 64), ((2, 1, 3, 48, 16), False, 64),
                                                  ((1, 1, 1, 16, 16), True, 64), ((1, 1, 5, 32, 32), True, 128)])
def test_fused_head_matches_the_unfused_head(half, shape, learn_sigma, ch):
    """out.0 GroupNorm -> SiLU -> out.2 conv (unet.py:993-997, 1043-1044): the single tcgen05 kernel (GroupNorm affine
    applied while staging, contraction over channels once per voxel, taps as a shifted sum; fp16 operands) against
    the GroupNorm pass + fp32 CUDA-core head.  Ragged tiles, z-ranges with clipped halos, batch 2, one output channel."""
    over = dict(large_size=16, small_size=16, num_channels=ch, num_res_blocks=1, num_head_channels=64,
                learn_sigma=learn_sigma, timestep_respacing="10")
    low, x, _ = synth_inputs(shape, 0)
    t = torch.tensor([400] * shape[0], device=DEV)
    outs = {}
    for fused in (1, 0):
        model, _, _, _ = build(over, seed=12, fp16=half)
        model.set_option("head_tc", fused)
        model.set_option("profile", 1)
        outs[fused] = model(x.to(DEV), t, low_res=low.to(DEV)).cpu()
        n_small = [k for k, _, _ in model.profile_read()].count("conv_small")
        assert n_small == 2  # stem + head, either way
    err = max_rel(outs[1], outs[0])
    print(f"fused vs unfused head, mode {half}, shape {shape}: max-rel {err:.2e}")
    assert err <= 1e-3  # fp16 rounding of hn and of the head weights, fp32 accumulation


@pytest.mark.parametrize("opts", [dict(strip=2, fold_identity=1), dict(strip=1), dict(strip=0), dict(fold_identity=0), dict(cluster=1), dict(stem_tc=0),
                                  dict(head_tc=0), dict(pdl=0), dict(pdl=1)])
def test_c2_architecture_with_kernel_options(opts):
    """The optional kernel variants (strip staging with swapped MMA operands, identity skips folded into the
    accumulation as unit-weight 1x1x1 sources, weight multicast in 2-CTA clusters) on the shipped network:
    same parity bound against the CPU oracle as the default path."""
    from oracle.unet import unet_forward
    from oracle.weights import synth_state_dict
    from ddpm3d_b200 import script_util as su
    flags = cases.sr_flags(use_fp16=True)
    cfg = cases.cfg_from_flags(flags)
    sd = synth_state_dict(cfg, seed=4)
    model, _ = su.sr_create_model_and_diffusion(**flags)
    model.load_state_dict(sd)
    model.to(DEV)
    model.convert_to_fp16()
    model.eval()
    for k, v in opts.items():
        model.set_option(k, v)
    shape = (1, 1, 8, 96, 96)
    low, x, _ = synth_inputs(shape, 0)
    t = torch.tensor([777])
    want = unet_forward(cfg, sd, x, t, low)
    out = model(x.to(DEV), t.to(DEV), low_res=low.to(DEV)).cpu()
    err = max_rel(out, want)
    print(f"C2 architecture, options {opts}: eps max-rel {err:.3e}")
    assert err <= TOL[True], err


def test_operand_descriptor_may_start_at_any_row_of_a_swizzled_tile():
    """The hardware property the strip kernel relies on: a SWIZZLE_128B K-major operand descriptor whose start
    address is any multiple of 128 bytes reads the rows from there on (the swizzle is a function of absolute smem
    address bits; the descriptor's base-offset field stays 0).  D = A[shift : shift + 128] @ I must be exact."""
    g = torch.Generator().manual_seed(0)
    A = torch.randn((256, 64), generator=g).bfloat16().to(DEV)
    ident = torch.eye(64).bfloat16().to(DEV)
    out = torch.empty((128, 64), device=DEV)
    from gpu_util import stream
    for shift in (0, 1, 3, 7, 8, 13, 50, 99, 128):
        N.check(N.lib().ddpm3d_k_probe_rowshift(N.ptr(A), 256, N.ptr(ident), shift, 0, N.ptr(out), stream()))
        torch.cuda.synchronize()
        assert torch.equal(out, A[shift:shift + 128].float()), shift


@pytest.mark.parametrize("dt", [N.BF16, N.FP16])
@pytest.mark.parametrize("Cout,shape", [(32, (1, 4, 8, 8)), (128, (2, 3, 10, 7)), (256, (1, 2, 16, 16)), (96, (1, 5, 9, 11)),
                                        (128, (1, 6, 24, 24))])
def test_stem_tensor_core_tile(dt, Cout, shape):
    """input_blocks.0.0 (unet.py:809-811): Conv3d(2 -> C) as one M128 x C x K64 tcgen05 tile per 128 voxels (hand-swizzled
    im2col rows) vs F.conv3d fp32 and vs the CUDA-core stem kernel; ragged last tile, B > 1, every TMEM width."""
    B, Z, H, W = shape
    tdt = TDT[dt]
    g = torch.Generator().manual_seed(Cout + Z)
    x = torch.randn((B, 2, Z, H, W), generator=g).to(tdt).float()
    w = (torch.randn((Cout, 2, 3, 3, 3), generator=g) / np.sqrt(54)).to(tdt).float()
    b = torch.randn(Cout, generator=g)
    ref = F.conv3d(x, w, b, padding=1)
    args = (to_cl(x, tdt), pack_weight(w, tdt), b.to(DEV), None, B, Z, H, W, 2, Cout)
    tc = conv3d(dt, 3, *args)
    cc = conv3d(dt, 4, *args)
    assert max_rel(from_cl(tc), ref) <= ROUND_TOL[dt]
    assert max_rel(from_cl(cc), ref) <= ROUND_TOL[dt]
    # same products, fp32 accumulation in a different order, one rounding each: at most an ulp apart
    assert max_rel(from_cl(tc), from_cl(cc)) <= ROUND_TOL[dt]
