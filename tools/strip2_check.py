"""The shipped network on the full 96^3 patch with the strip variant restricted to the large layers (strip = 1) and with
plane-pair strips on the 12 x 12 level (strip = 2): output difference (summation order + one 16-bit rounding per tensor)
and graph-replayed evaluation time, alternated in one process."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ddpm3d_b200 import script_util as su

dev = torch.device("cuda", 0)
model, diffusion = su.sr_create_model_and_diffusion(**bench.C2_FLAGS)
model.load_state_dict(bench.synth_weights(model._specs))
model.to(dev); model.convert_to_fp16(); model.eval()
g = torch.Generator().manual_seed(0)
x = torch.randn(bench.PATCH, generator=g).to(dev); low = torch.rand(bench.PATCH, generator=g).to(dev)
t = torch.tensor([500.0], device=dev)
outs = {}
for lvl in (1, 2):
    model.set_option("strip", lvl)
    a = model(x, t, low_res=low).float().cpu()
    b = model(x, t, low_res=low).float().cpu()
    assert torch.equal(a, b), "not bit-identical from run to run"
    outs[lvl] = a
d = (outs[2] - outs[1]).abs().max() / outs[1].abs().max()
print(f"strip=2 vs strip=1 on the full C2 evaluation: max-rel {float(d):.3e}")
