"""-m gpu: the "library bar" of SURVEY.md section 8(d) -- what the reference itself would do on this GPU: the same
network as PyTorch eager ops (cuDNN convolutions, torch GroupNorm / SiLU / pooling) in the reference's use_fp16 flow,
here through the oracle's functional restatement of unet.py (the reference tree does not exist on the GPU box).
The native path must agree with it and must not be slower; the measured ratio is printed for DESIGN.md."""
import pytest
import torch

from ddpm3d_b200 import script_util as su
from oracle import cases
from oracle.unet import unet_forward
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu

from gpu_util import DEV, max_rel  # noqa: E402


def _time(fn, warm, reps):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


@pytest.mark.parametrize("half", ["fp16", "bf16"])
def test_native_path_vs_torch_cudnn_eager(half):
    """C2 (shipped architecture, one FULL 96^3 patch): UNet evaluation, native kernels vs torch-eager + cuDNN.
    Speed: against the eager 16-bit flow (what the reference would execute here).  Parity: against the same eager
    network in fp32 (TF32 off) -- the full-size counterpart of the oracle comparison on the 8-plane slab."""
    tdt = {"fp16": torch.float16, "bf16": torch.bfloat16}[half]
    flags = cases.sr_flags(use_fp16=True)
    cfg = cases.cfg_from_flags(flags)
    sd = synth_state_dict(cfg, seed=4)
    model, _ = su.sr_create_model_and_diffusion(**flags)
    model.load_state_dict(sd)
    model.to(DEV)
    model.set_half_dtype(half)
    model.convert_to_fp16()
    model.eval()
    # fp16_util.py:15-22 convert_module_to_f16: conv weights / biases of the torso only
    torso = ("input_blocks.", "middle_block.", "output_blocks.")
    conv_leaf = ("in_layers.2.", "out_layers.3.", "skip_connection.", ".op.", ".conv.", "qkv.", "proj_out.", "input_blocks.0.0.")
    sd_dev, sd32 = {}, {}
    for k, v in sd.items():
        is_conv = k.startswith(torso) and any(c in k for c in conv_leaf)
        sd_dev[k] = v.to(DEV, tdt if is_conv else torch.float32)
        sd32[k] = v.to(DEV)
    g = torch.Generator().manual_seed(0)
    shape = (1, 1, 96, 96, 96)
    x = torch.randn(shape, generator=g).to(DEV)
    low = torch.rand(shape, generator=g).to(DEV)
    t = torch.tensor([500], device=DEV)
    torch.backends.cudnn.benchmark = True
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            want32 = unet_forward(cfg, sd32, x, t, low).cpu()
            ms_lib, lib16 = _time(lambda: unet_forward(cfg, sd_dev, x, t, low, dtype=tdt), 2, 3)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    ms_nat, got = _time(lambda: model(x, t, low_res=low), 3, 10)
    err = max_rel(got.cpu(), want32)
    err_lib = max_rel(lib16.float().cpu(), want32)
    print(f"\nC2 UNet evaluation on the full 96^3 patch, {half}: native {ms_nat:.2f} ms, torch-eager + cuDNN {ms_lib:.2f} ms "
          f"({ms_lib / ms_nat:.2f}x); eps max-rel vs eager fp32: native {err:.2e}, eager {half} {err_lib:.2e}")
    assert err <= (1e-2 if half == "bf16" else 5e-3), err  # the north star's bf16 bound, at full size
    assert ms_nat <= ms_lib
