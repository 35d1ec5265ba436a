"""A/B a library option inside one process (alternating, to cancel box / clock drift):
python tools/ab_option.py <option>[:a:b] [reps] [fixed=value ...]  -> per-kind ms for option=a and option=b (default 1 and 0)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ddpm3d_b200 import script_util as su

opt = sys.argv[1]
VALS = (1, 0)
if ":" in opt:
    opt, va, vb = opt.split(":")
    VALS = (int(va), int(vb))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fixed = [a.split("=") for a in sys.argv[3:]]
dev = torch.device("cuda", 0)
model, diffusion = su.sr_create_model_and_diffusion(**bench.C2_FLAGS)
model.load_state_dict(bench.synth_weights(model._specs))
model.to(dev); model.convert_to_fp16(); model.eval()
g = torch.Generator().manual_seed(0)
x = torch.randn(bench.PATCH, generator=g).to(dev); low = torch.rand(bench.PATCH, generator=g).to(dev)
t = torch.tensor([500.0], device=dev)
for k_, v_ in fixed:
    model.set_option(k_, int(v_))
for _ in range(3):
    model(x, t, low_res=low)
res = {v: collections.defaultdict(float) for v in VALS}
wall = {v: [] for v in VALS}
for r in range(reps):
    for v in VALS:
        model.set_option(opt, v)
        model.set_option("profile", 0)
        for _ in range(2):
            model(x, t, low_res=low)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            model(x, t, low_res=low)
        e1.record(); torch.cuda.synchronize()
        wall[v].append(e0.elapsed_time(e1) / 5)
        model.set_option("profile", 1)
        model(x, t, low_res=low); model.profile_read()
        model(x, t, low_res=low)
        for kind, ms, work in model.profile_read():
            res[v][kind] += ms / reps
for v in VALS:
    print(f"{opt}={v}: graph-replayed eval {sorted(wall[v])[len(wall[v]) // 2]:.3f} ms (median of {reps}); per kind:",
          {k: round(m, 3) for k, m in res[v].items()})
