"""Per-launch table of one C2 UNet evaluation from the library's own profile mode
(CUDA events around every launch on the launching stream)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ddpm3d_b200 import script_util as su

dev = torch.device("cuda", 0)
shape = bench.PATCH
opts = [a for a in sys.argv[1:] if "=" in a]
pos = [a for a in sys.argv[1:] if "=" not in a and not a.startswith("--")]
if pos:
    z, h, w = (int(v) for v in pos[0].split(","))
    shape = (1, 1, z, h, w)
model, diffusion = su.sr_create_model_and_diffusion(**bench.C2_FLAGS)
model.load_state_dict(bench.synth_weights(model._specs))
model.to(dev); model.convert_to_fp16(); model.eval()
for o in opts:
    k, v = o.split("=")
    model.set_option(k, int(v))
g = torch.Generator().manual_seed(0)
x = torch.randn(shape, generator=g).to(dev); low = torch.rand(shape, generator=g).to(dev)
t = torch.tensor([500.0], device=dev)
for _ in range(2):
    model(x, t, low_res=low)
# profile = 2: the evaluation is captured as usual, the event pairs are nodes of the graph; the records are those of the
# last replay (profile = 1, eager launches, adds ~8 us of launch gap to every short kernel)
mode = 1 if "--eager" in sys.argv else 2
out = torch.empty((1, 2) + tuple(shape[2:]), device=dev)
model.set_option("profile", mode)
if mode == 2:
    import ctypes as C
    from ddpm3d_b200 import _native as N
    def fwd():  # fixed output buffer: one graph key
        N.check(N.lib().ddpm3d_unet_forward(model._ensure_ctx(), N.ptr(x), N.ptr(low), N.ptr(t), None, N.ptr(out), 1, shape[2], shape[3],
                                            shape[4], N.current_stream_ptr(dev)))
    for _ in range(4):
        fwd()
else:
    model(x, t, low_res=low)
    model.profile_read()
    model(x, t, low_res=low)
recs = model.profile_read()
tot = sum(r[1] for r in recs)
print(f"total {tot:.3f} ms over {len(recs)} launches")
agg = {}
for i, (kind, ms, work) in enumerate(recs):
    if kind.startswith("conv") or kind == "attention":
        rate = f"{work / ms / 1e9:8.1f} TFLOP/s" if ms > 0 else ""
    else:
        rate = f"{work / ms / 1e6:8.1f} GB/s" if ms > 0 and work > 0 else ""
    print(f"{i:3d} {kind:13s} {ms:8.4f} ms  work {work:.4g} {rate}")
