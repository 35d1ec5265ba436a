"""CPU oracle for the 3D-DDPM sampling hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy fp64 for the schedules and integer
respacing, torch-CPU fp32 functional ops for the UNet) of the reference's
algorithm for the path named in BASELINE.json `north_star`:

    SpacedDiffusion.p_sample_loop -> p_sample -> p_mean_variance -> q_posterior
    with SuperResModel_noatt (UNetModel_noatt) evaluated at every step.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it, and only as the checker / the timed CPU
baseline.  The product path (the package `3d-denoising-diffusion-model_b200`)
never imports it and has no CPU fallback.

Pinning: the reference ships no tests or golden vectors ("parity unpinned" by
the reference itself).  The oracle is instead pinned
against OUTPUTS OF THE REFERENCE ITSELF, imported unmodified from
/root/reference in the build container by `oracle/make_golden.py`; the
resulting fixtures live in `tests/golden/` and `tests/test_oracle_golden.py`
checks every oracle function against them (bit-exact for integer / fp64
schedule work, <=1e-5 max-rel for fp32 network outputs, where the only
difference is op order inside torch itself).

Every function cites the reference file:line it restates (paths relative to
/root/reference).
"""

from .schedule import (  # noqa: F401
    named_beta_schedule,
    space_timesteps,
    DiffusionTables,
    make_tables,
)
from .unet import UNetConfig, build_plan, param_specs, unet_forward, timestep_embedding  # noqa: F401
from .sampler import ddim_sample, p_mean_variance, p_sample, p_sample_loop  # noqa: F401
from .weights import synth_state_dict, synth_inputs  # noqa: F401
