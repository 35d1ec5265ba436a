"""Property tests (hypothesis) of the host logic against the oracle restatement, over inputs the fixed fixtures do
not reach: respacing specs, schedule tables, patch tiling, rank striding, slab bounds.  CPU only."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from ddpm3d_b200 import dist_util, script_util as su, volume
from ddpm3d_b200.respace import space_timesteps
from ddpm3d_b200.slab import slab_bounds
from oracle import schedule as osch, volume as ovol

FAST = settings(max_examples=60, deadline=None)


@st.composite
def spacing_specs(draw):
    T = draw(st.integers(2, 1200))
    kind = draw(st.sampled_from(["count", "ddim", "sections"]))
    if kind == "count":
        return T, str(draw(st.integers(1, T)))
    if kind == "ddim":
        return T, "ddim" + str(draw(st.integers(1, T)))
    k = draw(st.integers(1, 4))
    return T, ",".join(str(draw(st.integers(1, max(1, T // k)))) for _ in range(k))


@FAST
@given(spacing_specs())
def test_space_timesteps_equals_oracle(spec):
    """respace.py:7-60: same set, or the same refusal (a ddimN without an integer stride, a section too small)."""
    T, s = spec
    try:
        want = osch.space_timesteps(T, s)
    except ValueError:
        with pytest.raises(ValueError):
            space_timesteps(T, s)
        return
    got = space_timesteps(T, s)
    assert got == want and all(0 <= i < T for i in got)


@pytest.mark.filterwarnings("ignore::RuntimeWarning")  # 20-step linear schedules end at beta = 1: 1/alphas_cumprod = inf
@settings(max_examples=25, deadline=None)
@given(st.integers(20, 400), st.sampled_from(["linear", "cosine"]), st.booleans(), st.booleans(), st.data())
def test_respaced_tables_equal_oracle(steps, schedule, learn_sigma, sigma_small, data):
    """gaussian_diffusion.py:118-169 + respace.py:72-86 in fp64: every table bit-equal to the oracle's.  (Domain: the
    reference itself rejects linear schedules shorter than 20 steps -- beta > 1 -- and one-step processes.)"""
    n = data.draw(st.integers(2, steps))
    kw = dict(steps=steps, noise_schedule=schedule, learn_sigma=learn_sigma, sigma_small=sigma_small,
              timestep_respacing=str(n))
    d = su.create_gaussian_diffusion(**kw)
    t = osch.make_tables(**kw)
    assert list(d.timestep_map) == list(t.timestep_map) and d.num_timesteps == n
    for name in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
                 "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                 "posterior_mean_coef1", "posterior_mean_coef2"):
        a, b = np.asarray(getattr(d, name), dtype=np.float64), np.asarray(getattr(t, name), dtype=np.float64)
        assert a.shape == b.shape and np.array_equal(a.view(np.int64), b.view(np.int64)), name
    tab = d.step_scalars()
    assert len(tab) == n and tab[0].model_t == float(d.timestep_map[0])


@FAST
@given(st.integers(1, 400), st.integers(1, 400), st.integers(1, 400), st.sampled_from([8, 16, 96]))
def test_patch_grid_equals_oracle_and_covers_the_volume(D, H, W, P):
    """scripts/test.py:205-230, 280-299: same origins in the same order as the oracle; when the three fixed
    in-plane starts can cover a dimension (dim <= 3P) every voxel lies in some patch."""
    if H < P or W < P:
        H, W = max(H, P), max(W, P)   # the reference assumes in-plane dims of at least one patch
    want = [(z0, x0, y0) for x0 in ovol.xy_starts(H, P) for y0 in ovol.xy_starts(W, P) for z0 in ovol.z_starts(D, P)]
    got = volume.patch_grid(D, H, W, P)
    assert got == want and len(got) == 9 * (1 if D <= P else 2)
    for (z0, x0, y0) in got:
        assert 0 <= x0 <= H - P and 0 <= y0 <= W - P and 0 <= z0 <= max(D - P, 0)
    if H <= 3 * P and W <= 3 * P and D <= 2 * P:
        cov = np.zeros((D, H, W), dtype=bool)
        for (z0, x0, y0) in got:
            cov[z0:z0 + P, x0:x0 + P, y0:y0 + P] = True
        assert cov.all()


@FAST
@given(st.integers(0, 200), st.integers(1, 16))
def test_patch_indices_partition(n, world):
    """scripts/test.py:243 rank striding: disjoint, complete, balanced to within one patch."""
    parts = [dist_util.patch_indices(n, r, world) for r in range(world)]
    assert sorted(i for p in parts for i in p) == list(range(n))
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


@FAST
@given(st.integers(1, 2000), st.integers(1, 16))
def test_slab_bounds_partition(z, world):
    if z < world:
        with pytest.raises(ValueError):
            slab_bounds(z, world)
        return
    b = slab_bounds(z, world)
    sizes = np.diff(b)
    assert b[0] == 0 and b[-1] == z and len(b) == world + 1 and sizes.min() >= 1 and sizes.max() - sizes.min() <= 1
    assert list(sizes) == sorted(sizes, reverse=True)   # the longer slabs come first


@settings(max_examples=300, deadline=None)
@given(B=st.integers(1, 2), Z=st.integers(1, 200), H=st.integers(2, 200), W=st.integers(2, 200),
       cin=st.sampled_from([64, 128, 192, 256, 384, 512, 768]), cout=st.sampled_from([64, 128, 256, 384, 512]),
       taps=st.sampled_from([27, 9, 1]), extra=st.sampled_from([0, 128, 640]), split_k=st.integers(0, 1),
       strip=st.integers(0, 2), sms=st.sampled_from([148, 132, 80]))
def test_conv_planner_invariants(B, Z, H, W, cin, cout, taps, extra, split_k, strip, sms):
    """The conv planner (csrc/conv_tc.cu through ddpm3d_k_conv_plan; host arithmetic, no device) over arbitrary layer
    shapes: every plan is launchable (tile counts, MMA N, TMA box and shared-memory limits) and the options gate what
    they say they gate."""
    import ctypes as C
    from ddpm3d_b200 import _native as N
    out = (C.c_int32 * 8)()
    N.check(N.lib().ddpm3d_k_conv_plan(N.BF16, B, Z, H, W, cin, cout, taps, extra, split_k, strip, sms, out))
    kind, a, b, tiles, grid, c, steps, rows = list(out)
    assert kind in (1, 2, 3)                       # every channel count above is a multiple of 64: always eligible
    assert tiles >= 1 and 1 <= grid <= sms
    if kind == 3:                                  # strip kernel
        NP, NV, NW = a, b, c
        assert strip >= 1 and taps in (27, 9) and cout % 128 == 0
        assert NP in (1, 2) and NV % 16 == 0 and 160 <= NV <= 256      # tcgen05 N for M = 128; N = 256 fills the accumulator
        assert 3 <= NW <= 8 and rows <= 256                            # weight ring, TMA box extent
        assert grid == min(tiles, sms)
        if NP == 2:                                # plane pairs: level 2 only, one wave (no double-buffered accumulator)
            assert strip == 2 and tiles <= sms
        else:
            assert tiles >= 2 * sms
        # a tile covers NV consecutive positions of a padded band: the tiles of one (batch, plane group) cover it
        per_plane_group = tiles // (B * -(-Z // NP) * (cout // 128))
        assert per_plane_group * B * -(-Z // NP) * (cout // 128) == tiles
        assert per_plane_group * NV >= H * W       # (bands x tiles per band) x NV >= the plane, pad columns aside
        macro_main = (3 if taps == 27 else 1) * (cin // 64)
        assert steps == macro_main + extra // 64
    else:                                          # brick kernel
        MT, BN = a, b
        assert MT in (1, 2) and BN in (64, 128, 256) and cout % BN == 0 and MT * BN <= 256
        assert steps == taps * (cin // 64) + extra // 64
        if kind == 2:
            assert split_k == 1 and grid == sms and c >= 2      # stream-K: every SM gets an equal share of the k-steps
        else:
            assert grid == min(tiles, sms)
    if strip == 0:
        assert kind != 3
