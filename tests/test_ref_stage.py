"""CPU: the staged copy of the unmodified reference (oracle/_ref, built by `python -m oracle.build_ref` /
__graft_entry__.build() where /root/reference exists) is byte-identical to its manifest, imports, and agrees with the
oracle's restatement -- so `bench.py --impl reference` and the cpu_baseline / library_bar legs time the reference itself."""
import pytest
import torch

from oracle import build_ref, cases
from oracle.unet import param_specs, unet_forward
from oracle.weights import synth_inputs, synth_state_dict

pytestmark = pytest.mark.skipif(not build_ref.available(), reason="oracle/_ref not staged (no reference checkout)")


def test_staged_reference_is_unmodified():
    assert build_ref.verify()


def test_staged_reference_matches_the_oracle():
    su = build_ref.load_ref()
    case = cases.UNET_CASES["tiny"]
    flags = cases.sr_flags(**case["flags"])
    cfg = cases.cfg_from_flags(flags)
    model, diffusion = su.sr_create_model_and_diffusion(**flags)
    assert [(k, tuple(v.shape)) for k, v in model.state_dict().items()] == [(k, tuple(s)) for k, s in param_specs(cfg)]
    sd = synth_state_dict(cfg, seed=0)
    model.load_state_dict(sd)
    model.eval()
    low, x, _ = synth_inputs(case["shape"], 0)
    t = torch.tensor(case["t"])
    with torch.no_grad():
        want = model(x, t, low_res=low)
    got = unet_forward(cfg, sd, x, t, low)
    assert float((got - want).abs().max() / want.abs().max()) <= 1e-5
    assert diffusion.num_timesteps == 10
