"""Host mirror of the reference's live model class: SuperResModel_noatt / UNetModel_noatt
(guided_diffusion/unet.py:720-1044, 1676-1694).  It owns the fp32 state_dict (same keys and
shapes as the reference checkpoint) and a libddpm3d context; every operator of the forward pass
is a CUDA kernel inside that library.  There is no torch implementation of the network here and
no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict

import numpy as np

from . import _native as N

NUM_CLASSES = 1000  # script_util.py:8


def _make_config(*, image_size, model_channels, out_channels, num_res_blocks, attention_resolutions, channel_mult,
                 num_classes, num_heads, num_head_channels, num_heads_upsample, use_scale_shift_norm,
                 resblock_updown, use_new_attention_order, precision, x_channels=1, dims=3, middle_attention=False,
                 unconditional=False) -> N.Config:
    cfg = N.Config()
    cfg.image_size = int(image_size)
    cfg.in_channels = int(x_channels)
    cfg.dims = int(dims)
    cfg.middle_attention = int(bool(middle_attention))
    cfg.unconditional = int(bool(unconditional))
    cfg.model_channels = int(model_channels)
    cfg.out_channels = int(out_channels)
    cfg.num_res_blocks = int(num_res_blocks)
    if len(channel_mult) > N.MAX_LEVELS:
        raise ValueError("too many resolution levels")
    cfg.n_levels = len(channel_mult)
    for i, m in enumerate(channel_mult):
        if int(m) != m:
            raise ValueError("channel_mult entries must be integers")
        cfg.channel_mult[i] = int(m)
    ds = list(attention_resolutions)
    if len(ds) > N.MAX_LEVELS:
        raise ValueError("too many attention resolutions")
    cfg.n_attention_ds = len(ds)
    for i, d in enumerate(ds):
        cfg.attention_ds[i] = int(d)
    cfg.num_classes = int(num_classes) if num_classes is not None else 0
    cfg.num_heads = int(num_heads)
    cfg.num_head_channels = int(num_head_channels)
    cfg.num_heads_upsample = int(num_heads_upsample)
    cfg.use_scale_shift_norm = int(bool(use_scale_shift_norm))
    cfg.resblock_updown = int(bool(resblock_updown))
    cfg.use_new_attention_order = int(bool(use_new_attention_order))
    cfg.precision = precision
    return cfg


def timestep_freqs(dim, max_period=10000):
    """nn.py:113-115: the sinusoid frequencies, evaluated on the host in fp32 with torch exactly as the
    reference does (it then moves the table to the device)."""
    import torch
    half = dim // 2
    return torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).contiguous()


_HALF = {"bf16": N.BF16, "fp16": N.FP16, "bf16_strict": N.BF16_STRICT}


def _half_from_env():
    import os
    return _HALF[os.environ.get("DDPM3D_HALF_DTYPE", "bf16").lower()]


class _Ctx:
    """RAII wrapper of ddpm3d_ctx*."""

    def __init__(self, cfg: N.Config):
        self._h = C.c_void_p()
        N.check(N.lib().ddpm3d_create(C.byref(cfg), C.byref(self._h)))

    @property
    def _as_parameter_(self):
        return self._h

    def close(self):
        if self._h:
            N.lib().ddpm3d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def param_specs(self):
        L = N.lib()
        out = []
        key = C.c_char_p()
        shape = (C.c_int64 * 8)()
        nd = C.c_int()
        for i in range(L.ddpm3d_param_count(self)):
            N.check(L.ddpm3d_param_info(self, i, C.byref(key), shape, C.byref(nd)))
            out.append((key.value.decode(), tuple(int(shape[d]) for d in range(nd.value))))
        return out


def sampler_only_context(diffusion, device):
    """A context that only carries a schedule table (for p_sample_update with a foreign model)."""
    import torch
    cfg = _make_config(image_size=8, model_channels=32, out_channels=1, num_res_blocks=1, attention_resolutions=(0,),
                       channel_mult=(1,), num_classes=None, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
                       use_scale_shift_norm=True, resblock_updown=True, use_new_attention_order=False, precision=N.FP32)
    ctx = _Ctx(cfg)
    tab = diffusion.step_scalars()
    with torch.cuda.device(device):
        N.check(N.lib().ddpm3d_set_schedule(ctx, tab, len(tab), diffusion.mean_code, diffusion.var_code))
    return ctx


class UNetModel_noatt:
    """guided_diffusion/unet.py:720-1044.  Constructor arguments are the reference's.

    The sibling classes differ only in two switches the library's topology builder understands:
    `_middle_attention` (UNetModel, unet.py:539-563) and `_concat_low_res` (the SuperRes* classes, which double
    `in_channels` and concatenate `low_res` in forward, unet.py:1654-1694).  dims=2 networks (Conv2d, (B,C,H,W)
    images) run through the same kernels as one-plane volumes."""

    _middle_attention = False
    _concat_low_res = False

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False):
        import torch
        if dims not in (2, 3):
            raise NotImplementedError("dims must be 2 or 3 (1-D networks are not built)")
        if not conv_resample:
            raise NotImplementedError("conv_resample=False is not reachable from script_util and not built")
        if self._concat_low_res and in_channels % 2:
            raise ValueError("a SuperRes model's in_channels counts x and low_res")
        self.dims = dims
        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = tuple(attention_resolutions)
        self.dropout = dropout  # no-op at inference
        self.channel_mult = tuple(channel_mult)
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint  # no-op without autograd
        self.dtype = torch.float16 if use_fp16 else torch.float32  # unet.py:787; the torso runs bf16 on tcgen05
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample
        self.use_scale_shift_norm = use_scale_shift_norm
        self.resblock_updown = resblock_updown
        self.use_new_attention_order = use_new_attention_order
        self.training = True
        self._half = _half_from_env()
        self._precision = self._half if use_fp16 else N.FP32
        self._device = torch.device("cpu")
        self._ctx = None
        self._ctx_sig = None
        self._bound = None
        self._slab = None
        self._options = {}
        # enumerate the state_dict contract from the native topology builder (works without a GPU)
        probe = _Ctx(self._config(N.FP32))
        self._specs = probe.param_specs()
        probe.close()
        self._state = self._initial_state()

    # ---- construction helpers ---------------------------------------------------------------------
    def _config(self, precision):
        return _make_config(image_size=self.image_size, model_channels=self.model_channels,
                            out_channels=self.out_channels, num_res_blocks=self.num_res_blocks,
                            attention_resolutions=self.attention_resolutions, channel_mult=self.channel_mult,
                            num_classes=self.num_classes, num_heads=self.num_heads,
                            num_head_channels=self.num_head_channels, num_heads_upsample=self.num_heads_upsample,
                            use_scale_shift_norm=self.use_scale_shift_norm, resblock_updown=self.resblock_updown,
                            use_new_attention_order=self.use_new_attention_order, precision=precision,
                            x_channels=self._x_channels, dims=self.dims, middle_attention=self._middle_attention,
                            unconditional=not self._concat_low_res)

    @property
    def _x_channels(self):
        """Channels of the `x` argument of forward (the SuperRes classes split in_channels between x and low_res)."""
        return self.in_channels // 2 if self._concat_low_res else self.in_channels

    @property
    def _fused_sampler(self):
        """UNet + posterior update in one library call: built for the live one-channel conditional 3-D model."""
        return self._concat_low_res and self.in_channels == 2 and self.dims == 3

    def _initial_state(self):
        """Same families of initial values as the reference constructor: uniform(+-1/sqrt(fan_in)) for
        conv / linear, ones / zeros for GroupNorm, N(0,1) for label_emb, and zero_module
        (nn.py:68-74) on out_layers.3, proj_out and out.2 (unet.py:210-212,294,996)."""
        import torch
        sd = OrderedDict()
        fan = {}
        for key, shape in self._specs:
            stem, leaf = key.rsplit(".", 1)
            is_norm = stem.endswith(("in_layers.0", "out_layers.0", ".norm")) or stem == "out.0"
            zero = stem.endswith(("out_layers.3", "proj_out")) or stem == "out.2"
            if is_norm:
                sd[key] = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
            elif zero:
                sd[key] = torch.zeros(shape)
            elif key == "label_emb.weight":
                sd[key] = torch.randn(shape)
            else:
                if leaf == "weight":
                    fan[stem] = math.prod(shape[1:])
                bound = 1.0 / math.sqrt(fan[stem])
                sd[key] = (torch.rand(shape) * 2 - 1) * bound
        return sd

    # ---- nn.Module surface used by scripts/test.py:29-35 ------------------------------------------
    def state_dict(self):
        return OrderedDict((k, v.clone()) for k, v in self._state.items())

    def load_state_dict(self, state_dict, strict=True):
        import torch
        expected = dict(self._specs)
        missing = [k for k in expected if k not in state_dict]
        unexpected = [k for k in state_dict if k not in expected]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing keys {missing}, unexpected keys {unexpected}")
        for k, v in state_dict.items():
            if k not in expected:
                continue
            if tuple(v.shape) != expected[k]:
                raise RuntimeError(f"size mismatch for {k}: checkpoint {tuple(v.shape)} vs model {expected[k]}")
            self._state[k] = v.detach().to("cpu", torch.float32).contiguous().clone()
        self._drop_ctx()
        return missing, unexpected

    def parameters(self):
        dev = self._device
        for v in self._state.values():
            yield v if dev.type == "cpu" else _DeviceTag(v, dev)

    def named_parameters(self):
        for k, v in self._state.items():
            yield k, v

    def to(self, device=None, *args, **kwargs):
        import torch
        if device is not None and not isinstance(device, torch.dtype):
            d = torch.device(device)
            if d.type == "cuda" and d.index is None:
                d = torch.device("cuda", torch.cuda.current_device())
            if d != self._device:
                self._device = d
                self._drop_ctx()
        return self

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("training is out of scope of the B200 sampling path")
        return self.eval()

    def requires_grad_(self, flag=False):
        return self

    def convert_to_fp16(self):
        """unet.py:999-1005.  The torso (input/middle/output blocks) runs in 16-bit on tcgen05 tensor cores
        (bf16 by default as BASELINE.json names it; the reference's own fp16 with set_half_dtype("fp16"));
        time_embed, emb_layers, GroupNorm statistics and the `out` head stay fp32 like the reference."""
        if self._precision != self._half:
            self._precision = self._half
            self._drop_ctx()

    def set_half_dtype(self, name):
        """Which 16-bit mode `use_fp16` / convert_to_fp16() means: "bf16" (default, or env DDPM3D_HALF_DTYPE): bf16
        tensor-core operands, block inputs / outputs stored as fp16; "fp16" (the reference's dtype everywhere: same
        tensor-core rate, 3 more mantissa bits on the operands too); "bf16_strict" (every 16-bit tensor bf16)."""
        half = _HALF[name]
        if half != self._half:
            was_half = self._precision in (N.BF16, N.FP16, N.BF16_STRICT)
            self._half = half
            if was_half:
                self._precision = half
                self._drop_ctx()
        return self

    def convert_to_fp32(self):
        """unet.py:1007-1013."""
        if self._precision != N.FP32:
            self._precision = N.FP32
            self._drop_ctx()

    # ---- native context -------------------------------------------------------------------------------
    def set_option(self, name, value):
        """Library knobs (include/ddpm3d.h ddpm3d_set_option): cuda_graph, conv_path, profile."""
        self._options[name] = int(value)
        if self._ctx is not None:
            N.check(N.lib().ddpm3d_set_option(self._ctx, name.encode(), int(value)))

    def _drop_ctx(self):
        if self._ctx is not None:
            self._ctx.close()
        self._ctx = None
        self._bound = None
        self._slab = None  # the NCCL communicator lived in the context

    def _ensure_ctx(self):
        import torch
        if self._ctx is not None:
            return self._ctx
        if self._device.type != "cuda":
            raise RuntimeError("the B200 path has no CPU fallback: move the model to a CUDA device first (model.to(dev))")
        ctx = _Ctx(self._config(self._precision))
        L = N.lib()
        for key, shape in self._specs:
            t = self._state[key]
            shp = (C.c_int64 * len(shape))(*shape)
            N.check(L.ddpm3d_load_tensor(ctx, key.encode(), N.ptr(t), shp, len(shape)))
        N.check(L.ddpm3d_finalize_weights(ctx, self._device.index))
        freqs = timestep_freqs(self.model_channels)
        N.check(L.ddpm3d_set_timestep_freqs(ctx, N.ptr(freqs), freqs.numel()))
        for k, v in self._options.items():
            N.check(L.ddpm3d_set_option(ctx, k.encode(), v))
        self._ctx = ctx
        return ctx

    def _bind_schedule(self, diffusion):
        import torch
        ctx = self._ensure_ctx()
        if self._bound is not diffusion:
            tab = diffusion.step_scalars()
            with torch.cuda.device(self._device):
                N.check(N.lib().ddpm3d_set_schedule(ctx, tab, len(tab), diffusion.mean_code, diffusion.var_code))
            self._bound = diffusion

    # ---- one large volume as z-slabs over the ranks of a process group ---------
    def enable_slab_sharding(self, group=None):
        """Creates the library's NCCL communicator over `group` (one process per GPU).  Afterwards forward /
        p_sample / the sampling loop take this rank's z-slab; call set_slab() with its position first."""
        import torch.distributed as dist
        ctx = self._ensure_ctx()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        buf = (C.c_char * 128)()
        if rank == 0:
            N.check(N.lib().ddpm3d_comm_unique_id(buf))
        obj = [bytes(buf) if rank == 0 else None]
        dist.broadcast_object_list(obj, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        idbuf = (C.c_char * 128).from_buffer_copy(obj[0])
        import torch
        with torch.cuda.device(self._device):
            N.check(N.lib().ddpm3d_set_comm(ctx, idbuf, rank, world))
        self._slab = (rank, world)
        return self

    def set_slab(self, z_begin, z_total):
        """Position of this rank's slab in the global volume; switches the sharded path on (every following call
        must be a proper slab of that volume)."""
        N.check(N.lib().ddpm3d_set_slab(self._ensure_ctx(), int(z_begin), int(z_total)))

    def disable_slab_sharding(self):
        """Back to un-sharded work on this context (independent patches, ensemble samples); the communicator stays."""
        if self._ctx is not None:
            N.check(N.lib().ddpm3d_set_slab(self._ctx, 0, 0))

    def launch_count(self):
        return int(N.lib().ddpm3d_launch_count(self._ctx)) if self._ctx is not None else 0

    def profile_read(self):
        """After running with set_option('profile', 1): list of (kind name, ms, work)."""
        L = N.lib()
        cap = 1 << 16
        buf = (N.ProfRecord * cap)()
        n = N.check(L.ddpm3d_profile_read(self._ctx, buf, cap))
        return [(N.PROF_KINDS[buf[i].kind], float(buf[i].ms), float(buf[i].work)) for i in range(min(n, cap))]

    def workspace_bytes(self, B, Z, H, W):
        ctx = self._ensure_ctx()
        return int(N.check(N.lib().ddpm3d_workspace_bytes(ctx, B, Z, H, W)))

    # ---- forward (unet.py:1015-1044) -------------------------------------------------------------------
    def _check_io(self, x, low_res):
        if x.dim() != self.dims + 2 or x.shape[1] != self._x_channels:
            raise AssertionError(f"x must be (B, {self._x_channels}, {'Z, ' if self.dims == 3 else ''}H, W)")
        if self._concat_low_res:
            if low_res is None or tuple(low_res.shape) != tuple(x.shape):
                raise AssertionError("low_res must have the shape of x (unet.py:1690-1693)")
        elif low_res is not None:
            raise TypeError("forward() got an unexpected keyword argument 'low_res'")  # unet.py:687 / :1015
        if x.device != self._device:
            raise RuntimeError(f"input on {x.device} but model on {self._device}")

    def _y(self, y, B):
        import torch
        if (y is not None) != (self.num_classes is not None):
            raise AssertionError("must specify y if and only if the model is class-conditional")  # unet.py:1024-1026
        if y is None:
            return None
        assert y.shape == (B,)
        return y.to(self._device, torch.int64).contiguous()

    def forward(self, x, timesteps, low_res=None, y=None):
        import torch
        ctx = self._ensure_ctx()
        self._check_io(x, low_res)
        B = x.shape[0]
        Z = x.shape[2] if self.dims == 3 else 1
        H, W = x.shape[-2:]
        x = x.contiguous().float()
        low = low_res.contiguous().float() if low_res is not None else None
        t = timesteps.to(self._device).float().contiguous()  # nn.py:117 casts to float anyway
        assert t.shape == (B,)
        yy = self._y(y, B)
        out = torch.empty((B, self.out_channels, *x.shape[2:]), device=self._device, dtype=torch.float32)
        with torch.cuda.device(self._device):
            N.check(N.lib().ddpm3d_unet_forward(ctx, N.ptr(x), N.ptr(low), N.ptr(t), N.ptr(yy), N.ptr(out),
                                                B, Z, H, W, N.current_stream_ptr(self._device)))
        return out

    __call__ = forward

    # ---- fused sampler entry points used by GaussianDiffusion ---------------------------------------
    def _p_sample(self, diffusion, x, noise, step_index, model_kwargs, clip_denoised, clone=False):
        """UNet + posterior update for one step in a single (graph-cached) library call.  `step_index` is a
        Python int (same step for the whole batch) or a device int64 tensor (B,) as in the public p_sample."""
        import torch
        self._bind_schedule(diffusion)
        low = model_kwargs.get("low_res")
        self._check_io(x, low)
        B, _, Z, H, W = x.shape
        low = low.contiguous().float()
        yy = self._y(model_kwargs.get("y"), B)
        if not hasattr(self, "_ps_buf") or self._ps_buf[0].shape != x.shape or self._ps_buf[0].device != x.device:
            self._ps_buf = [torch.empty_like(x) for _ in range(3)]
            self._ps_stage = [torch.empty_like(x) for _ in range(3)]
        if clone:
            # public p_sample: the caller's x / noise / low_res are usually fresh allocations every step, and the
            # library's CUDA-graph cache is keyed by pointers -- stage them in persistent buffers (three small
            # device copies) so one captured graph serves every call
            for buf, src in zip(self._ps_stage, (x, noise, low)):
                buf.copy_(src, non_blocking=True)
            x, noise, low = self._ps_stage
        # two sample buffers ping-pong (the input `x` is usually the previous output); x0 has its own
        sample = self._ps_buf[1] if x.data_ptr() == self._ps_buf[0].data_ptr() else self._ps_buf[0]
        x0 = self._ps_buf[2]
        self._keep = (low, yy, noise)
        with torch.cuda.device(self._device):
            if torch.is_tensor(step_index):
                # stable address for the graph cache: the indices are copied into a persistent device buffer
                if getattr(self, "_t_buf", None) is None or self._t_buf.shape[0] != B or self._t_buf.device != x.device:
                    self._t_buf = torch.zeros(B, dtype=torch.int64, device=self._device)
                self._t_buf.copy_(step_index.to(self._device, torch.int64), non_blocking=True)
                tt = self._t_buf
                self._keep = (low, yy, noise, tt)
                N.check(N.lib().ddpm3d_p_sample_t(self._ctx, N.ptr(x), N.ptr(low), N.ptr(yy), N.ptr(noise), N.ptr(tt),
                                                  int(bool(clip_denoised)), N.ptr(sample), N.ptr(x0), B, Z, H, W,
                                                  N.current_stream_ptr(self._device)))
            else:
                N.check(N.lib().ddpm3d_p_sample(self._ctx, N.ptr(x), N.ptr(low), N.ptr(yy), N.ptr(noise), int(step_index),
                                                int(bool(clip_denoised)), N.ptr(sample), N.ptr(x0), B, Z, H, W,
                                                N.current_stream_ptr(self._device)))
        if clone:
            return {"sample": sample.clone(), "pred_xstart": x0.clone()}
        return {"sample": sample, "pred_xstart": x0}

    def _sample_loop(self, diffusion, x_T, model_kwargs, step_noise, seed, clip_denoised, n_steps=0):
        import torch
        self._bind_schedule(diffusion)
        low = model_kwargs.get("low_res")
        self._check_io(x_T, low)
        B, _, Z, H, W = x_T.shape
        low = low.contiguous().float()
        yy = self._y(model_kwargs.get("y"), B)
        if step_noise is not None:
            step_noise = step_noise.to(self._device).contiguous().float()
            steps = n_steps if n_steps > 0 else diffusion.num_timesteps
            assert step_noise.numel() >= steps * x_T.numel(), "step_noise must hold one tensor per executed step"
        out = torch.empty_like(x_T)
        with torch.cuda.device(self._device):
            N.check(N.lib().ddpm3d_sample_loop(self._ctx, N.ptr(x_T), N.ptr(low), N.ptr(yy), N.ptr(step_noise),
                                               C.c_uint64(int(seed)), int(bool(clip_denoised)), int(n_steps),
                                               N.ptr(out), B, Z, H, W, N.current_stream_ptr(self._device)))
        return out


class _DeviceTag:
    """What `next(model.parameters()).device` (gaussian_diffusion.py:507-508) needs."""

    def __init__(self, t, device):
        self._t = t
        self.device = device
        self.shape = t.shape
        self.dtype = t.dtype

    def numel(self):
        return self._t.numel()


class SuperResModel_noatt(UNetModel_noatt):
    """unet.py:1676-1694: concatenates `low_res` on the channel axis (done inside the library's
    pack_input kernel, never materialised by torch)."""

    _concat_low_res = True

    def __init__(self, image_size, in_channels, *args, **kwargs):
        super().__init__(image_size, int(in_channels * 2), *args, **kwargs)


class UNetModel(UNetModel_noatt):
    """unet.py:396-716: the full UNet -- an AttentionBlock sits between the two middle ResBlocks (:539-563).
    create_model (script_util.py:130-184) instantiates it with dims=2 on RGB images."""

    _middle_attention = True


class SuperResModel(UNetModel):
    """unet.py:1654-1673: UNetModel with `low_res` concatenated as is (same shape as x; the bilinear up-sampling
    of the upstream code is commented out in the reference)."""

    _concat_low_res = True

    def __init__(self, image_size, in_channels, *args, **kwargs):
        super().__init__(image_size, int(in_channels * 2), *args, **kwargs)
