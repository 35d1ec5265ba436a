"""-m gpu, needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`): one volume sharded as z-slabs
over two ranks (halo exchange + GroupNorm all-gather + K/V all-gather for attention blocks, over NCCL) must
reproduce the single-GPU result."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from ddpm3d_b200 import script_util as su, slab
    from oracle import cases
    from oracle.weights import synth_inputs, synth_state_dict
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    res = {}
    try:
        # the last case is the shipped network on 96x96 planes: 8 planes per rank make 296 strip tiles, so the
        # halo planes are read by the strip kernel's row-shifted descriptors (the C4 bench configuration)
        # the "8,4" / "8" cases carry attention blocks (ds = 2, 4): local queries against all-gathered keys / values,
        # CUDA-core kernel in fp32, tcgen05 kernel (64-wide heads) in bf16, batch 2 = one all-gather per element
        pick = os.environ.get("DDPM3D_SLAB_CASES")  # e.g. "0,5": run a subset (quick re-checks)
        for idx, (mode, shape, tol, ch, att) in enumerate((
                                          (False, (1, 1, 10, 16, 16), 2e-5, 64, "1000"),
                                          (False, (2, 1, 9, 16, 32), 2e-5, 64, "1000"),
                                          (True, (1, 1, 12, 32, 32), 1e-2, 64, "1000"),
                                          (False, (1, 1, 8, 16, 16), 2e-5, 64, "8,4"),
                                          (False, (2, 1, 4, 16, 16), 2e-5, 64, "8"),
                                          (True, (1, 1, 8, 32, 32), 1.3e-2, 64, "8,4"),
                                          (True, (1, 1, 16, 96, 96), 1e-2, 128, "1000"))):
            if pick and str(idx) not in pick.split(","):
                continue
            over = dict(large_size=16, small_size=16, num_channels=ch, num_res_blocks=2, num_head_channels=64,
                        timestep_respacing="10" if ch == 64 else "3", use_fp16=bool(mode), attention_resolutions=att)
            flags = cases.sr_flags(**over)
            cfg = cases.cfg_from_flags(flags)
            sd = synth_state_dict(cfg, seed=11)

            def make(half=mode):
                m, d = su.sr_create_model_and_diffusion(**dict(flags, use_fp16=bool(half)))
                m.load_state_dict(sd)
                m.to(dev)
                if half:
                    m.convert_to_fp16()
                return m.eval(), d

            single, diffusion = make()
            sharded, _ = make()
            if idx == 0:  # equal slabs, peer path switched off: NCCL send/recv + all-gather (the other equal-slab cases
                sharded.set_option("slab_p2p", 0)  # exchange through peer-mapped memory, the unequal one cannot)
            sharded.enable_slab_sharding()
            low, x_T, _ = synth_inputs(shape, 0)
            low, x_T = low.to(dev), x_T.to(dev)
            B, _, Z, H, W = shape
            # (a) one evaluation: the sharded result against the single-GPU network in FP32 (16-bit cases: the north
            # star's eps bound, with the attention toy network listed in tests/test_gpu_model.py at its own bound)
            t = torch.tensor([555.0] * B, device=dev)
            want = make(False)[0](x_T, t, low_res=low) if mode else single(x_T, t, low_res=low)
            bounds = slab.slab_bounds(Z, world)
            z0, z1 = bounds[rank], bounds[rank + 1]
            sharded.set_slab(z0, Z)
            part = sharded(x_T[:, :, z0:z1].contiguous(), t, low_res=low[:, :, z0:z1].contiguous())
            got = slab.gather_slabs(part, bounds)
            e1 = float((got - want).abs().max() / want.abs().max())
            # (b) the whole loop, philox noise (one global field)
            want_s = diffusion.p_sample_loop(single, shape, noise=x_T, model_kwargs={"low_res": low}, rng="philox", seed=5)
            got_s = slab.sample_volume_slabs(sharded, diffusion, low, noise=x_T, rng="philox", seed=5)
            e2 = float((got_s - want_s).pow(2).mean().sqrt() / want_s.pow(2).mean().sqrt())  # NRMSE of the volume
            res[str((mode, shape, att))] = (e1, e2, tol, sharded.launch_count())
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_slab_sharding_matches_single_gpu():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    out = dict(q.get(timeout=10) for _ in range(2))
    for rank in (0, 1):
        for key, (e1, e2, tol, launches) in out[rank].items():
            print(f"rank {rank} {key}: forward max-rel {e1:.2e}, loop NRMSE {e2:.2e}, launches {launches}")
            assert e1 <= tol, (key, e1)
            # fp32: only the GroupNorm summation order differs; bf16: two roundings-equivalent evaluations
            # of a 10-step loop (the C1 bf16-vs-fp32 loop NRMSE is 8e-3, tests/test_gpu_model.py)
            if "96, 96" in key:
                continue  # 3-step loop of the big case: only the forward parity is asserted
            assert e2 <= (1e-4 if tol < 1e-3 else 5e-2), (key, e2)
