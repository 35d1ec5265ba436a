"""Debug: bf16-mode error of the UNet vs the reference fixtures, SIMT vs tcgen05."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import cases
from oracle.weights import synth_inputs
from oracle.unet import unet_forward
from test_gpu_model import build
from gpu_util import DEV, max_rel
g = np.load(os.path.join(ROOT, "tests/golden/unet_tiny.npz"))
for name, case in cases.UNET_CASES.items():
    want = torch.from_numpy(g[f"{name}/out"])
    low, x, _ = synth_inputs(case["shape"], 0)
    kw = {"y": torch.tensor(case["y"], device=DEV)} if "y" in case else {}
    for fp16 in (False, True):
        for path in ((1,) if not fp16 else (1, 2)):
            model, _, cfg, sd = build(case["flags"], seed=case.get("seed", 0), fp16=fp16)
            model.set_option("conv_path", path)
            out = model(x.to(DEV), torch.tensor(case["t"], device=DEV), low_res=low.to(DEV), **kw).cpu()
            d = (out - want)
            print(f"{name:10s} fp16={fp16} path={path} max_rel={max_rel(out, want):.3e} rms_rel={float(d.pow(2).mean().sqrt()/want.pow(2).mean().sqrt()):.3e}", flush=True)
    # what does torch's own bf16 / fp16 autocast-free cast give on CPU? (precision floor of the dtype)
    sdh = {k: v for k, v in sd.items()}

# C1 loop in bf16
from oracle import cases as cs
gl = np.load(os.path.join(ROOT, "tests/golden/c1_loop.npz"))
for fp16 in (False, True):
    model, diffusion, _, _ = build(cs.C1_FLAGS, fp16=fp16)
    T = diffusion.num_timesteps
    low, x_T, noises = synth_inputs(cs.C1_SHAPE, T)
    s2 = diffusion.p_sample_loop(model, cs.C1_SHAPE, noise=x_T.to(DEV), clip_denoised=True, model_kwargs={"low_res": low.to(DEV)}, step_noise=torch.stack(noises).to(DEV)).cpu()
    want = torch.from_numpy(gl["sample"]); err = s2 - want
    print("C1 loop fp16=", fp16, "nrmse", float(err.pow(2).mean().sqrt()/want.pow(2).mean().sqrt()), "psnr", float(10*torch.log10(4.0/err.pow(2).mean())), "maxabs", float(err.abs().max()))
