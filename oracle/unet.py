"""Oracle: the conditional 3-D UNet (SuperResModel_noatt / UNetModel_noatt) as a
pure function of (config, state_dict, x, t, low_res).

TEST INFRASTRUCTURE (see oracle/__init__.py).  torch-CPU fp32 functional ops;
no nn.Module.  The block plan produced by `build_plan` is also the contract
the CUDA library's native topology builder is tested against (same state_dict
keys and shapes as the reference's).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class UNetConfig:
    """Arguments of UNetModel_noatt.__init__ (guided_diffusion/unet.py:751-772)
    as sr_create_model fills them (script_util.py:334-450)."""

    image_size: int = 96
    in_channels: int = 1  # doubled by SuperResModel_noatt (unet.py:1683-1685)
    model_channels: int = 128
    out_channels: int = 2
    num_res_blocks: int = 2
    attention_ds: tuple = (0,)
    channel_mult: tuple = (1, 1, 2, 3, 4)
    num_classes: Optional[int] = None
    num_heads: int = 4
    num_head_channels: int = 64
    num_heads_upsample: int = -1
    use_scale_shift_norm: bool = True
    resblock_updown: bool = True
    use_new_attention_order: bool = False
    conv_resample: bool = True
    # the other classes of unet.py: dims=2 Conv2d networks, UNetModel's middle attention (unet.py:539-563), and
    # models that take no low_res (UNetModel.forward, unet.py:687-716)
    dims: int = 3
    middle_attention: bool = False
    concat_low_res: bool = True

    @property
    def stem_cin(self) -> int:
        return self.in_channels * (2 if self.concat_low_res else 1)

    @staticmethod
    def from_model_flags(
        image_size, num_channels, num_res_blocks, channel_mult, learn_sigma, class_cond, attention_resolutions,
        num_heads, num_head_channels, num_heads_upsample, use_scale_shift_norm, resblock_updown,
        use_new_attention_order, **_ignored,
    ) -> "UNetConfig":
        """script_util.py:130-184 create_model: the 2-D RGB UNetModel."""
        if channel_mult == "":
            mult = {256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4)}[image_size]
        else:
            mult = tuple(int(m) for m in channel_mult.split(","))
        ds = tuple(image_size // int(r) for r in attention_resolutions.split(","))
        return UNetConfig(
            image_size=image_size, in_channels=3, model_channels=num_channels, out_channels=6 if learn_sigma else 3,
            num_res_blocks=num_res_blocks, attention_ds=ds, channel_mult=mult,
            num_classes=1000 if class_cond else None, num_heads=num_heads, num_head_channels=num_head_channels,
            num_heads_upsample=num_heads_upsample, use_scale_shift_norm=use_scale_shift_norm,
            resblock_updown=resblock_updown, use_new_attention_order=use_new_attention_order,
            dims=2, middle_attention=True, concat_low_res=False,
        )

    @staticmethod
    def from_sr_flags(
        large_size, num_channels, num_res_blocks, learn_sigma, class_cond,
        attention_resolutions, num_heads, num_head_channels, num_heads_upsample,
        use_scale_shift_norm, resblock_updown, **_ignored,
    ) -> "UNetConfig":
        """script_util.py:334-450 sr_create_model."""
        if large_size in (512, 256):
            mult = (1, 1, 2, 2, 4, 4)
        elif large_size == 64:
            mult = (1, 2, 3, 4)
        else:
            mult = (1, 1, 2, 3, 4)
        ds = tuple(large_size // int(r) for r in attention_resolutions.split(","))
        return UNetConfig(
            image_size=large_size,
            in_channels=1,
            model_channels=num_channels,
            out_channels=2 if learn_sigma else 1,
            num_res_blocks=num_res_blocks,
            attention_ds=ds,
            channel_mult=mult,
            num_classes=1000 if class_cond else None,
            num_heads=num_heads,
            num_head_channels=num_head_channels,
            num_heads_upsample=num_heads_upsample,
            use_scale_shift_norm=use_scale_shift_norm,
            resblock_updown=resblock_updown,
        )


# ----------------------------------------------------------------------------
# Block plan (unet.py:797-997)
# ----------------------------------------------------------------------------

def _res(prefix, cin, cout, up=False, down=False):
    return dict(kind="res", prefix=prefix, cin=cin, cout=cout, up=up, down=down)


def _attn(prefix, ch, heads):
    return dict(kind="attn", prefix=prefix, ch=ch, heads=heads)


def _heads(cfg: UNetConfig, ch: int, n: int) -> int:
    """unet.py:275-282."""
    if cfg.num_head_channels == -1:
        return n
    assert ch % cfg.num_head_channels == 0
    return ch // cfg.num_head_channels


def build_plan(cfg: UNetConfig) -> dict:
    """Replays the constructor of UNetModel_noatt (unet.py:797-997) and returns
    {'input': [[layer,...],...], 'middle': [...], 'output': [[...],...], 'out_norm_ch', 'out_conv_in'}."""
    mc = cfg.model_channels
    heads_up = cfg.num_heads if cfg.num_heads_upsample == -1 else cfg.num_heads_upsample
    ch = input_ch = int(cfg.channel_mult[0] * mc)
    inputs = [[dict(kind="conv", prefix="input_blocks.0.0", cin=cfg.stem_cin, cout=ch, stride_hw=1)]]
    chans = [ch]
    ds = 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            n = len(inputs)
            cout = int(mult * mc)
            layers = [_res(f"input_blocks.{n}.0", ch, cout)]
            ch = cout
            if ds in cfg.attention_ds:
                layers.append(_attn(f"input_blocks.{n}.1", ch, _heads(cfg, ch, cfg.num_heads)))
            inputs.append(layers)
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            n = len(inputs)
            if cfg.resblock_updown:
                inputs.append([_res(f"input_blocks.{n}.0", ch, ch, down=True)])
            else:
                # Downsample(ch, conv_resample=True, dims=3): strided conv (unet.py:129-133)
                inputs.append([dict(kind="conv", prefix=f"input_blocks.{n}.0.op", cin=ch, cout=ch, stride_hw=2)])
            chans.append(ch)
            ds *= 2
    if cfg.middle_attention:
        middle = [_res("middle_block.0", ch, ch), _attn("middle_block.1", ch, _heads(cfg, ch, cfg.num_heads)),
                  _res("middle_block.2", ch, ch)]
    else:
        middle = [_res("middle_block.0", ch, ch), _res("middle_block.1", ch, ch)]
    outputs = []
    outch = ch
    for level, _mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            inch = chans.pop()
            outch = chans.pop() if chans else inch
            n = len(outputs)
            layers = [_res(f"output_blocks.{n}.0", inch * 2, outch)]
            if ds in cfg.attention_ds:
                layers.append(_attn(f"output_blocks.{n}.{len(layers)}", outch, _heads(cfg, outch, heads_up)))
            if level and i == cfg.num_res_blocks:
                k = len(layers)
                if cfg.resblock_updown:
                    layers.append(_res(f"output_blocks.{n}.{k}", outch, outch, up=True))
                else:
                    layers.append(dict(kind="upconv", prefix=f"output_blocks.{n}.{k}.conv", cin=outch, cout=outch))
                ds //= 2
            outputs.append(layers)
            chans.append(outch)
    return dict(input=inputs, middle=middle, output=outputs, out_norm_ch=outch,
                out_conv_in=input_ch, time_embed_dim=mc * 4)


def param_specs(cfg: UNetConfig) -> list:
    """(state_dict key, shape) for every parameter, in the reference's
    registration order (time_embed, [label_emb], input_blocks, middle_block,
    output_blocks, out)."""
    plan = build_plan(cfg)
    mc, ted = cfg.model_channels, plan["time_embed_dim"]
    specs = [
        ("time_embed.0.weight", (ted, mc)), ("time_embed.0.bias", (ted,)),
        ("time_embed.2.weight", (ted, ted)), ("time_embed.2.bias", (ted,)),
    ]
    if cfg.num_classes is not None:
        specs.append(("label_emb.weight", (cfg.num_classes, ted)))

    def k(n):
        return (n,) * cfg.dims

    def layer_specs(L):
        p = L["prefix"]
        if L["kind"] in ("conv", "upconv"):
            return [(p + ".weight", (L["cout"], L["cin"], *k(3))), (p + ".bias", (L["cout"],))]
        if L["kind"] == "res":
            ci, co = L["cin"], L["cout"]
            e = 2 * co if cfg.use_scale_shift_norm else co
            out = [
                (p + ".in_layers.0.weight", (ci,)), (p + ".in_layers.0.bias", (ci,)),
                (p + ".in_layers.2.weight", (co, ci, *k(3))), (p + ".in_layers.2.bias", (co,)),
                (p + ".emb_layers.1.weight", (e, ted)), (p + ".emb_layers.1.bias", (e,)),
                (p + ".out_layers.0.weight", (co,)), (p + ".out_layers.0.bias", (co,)),
                (p + ".out_layers.3.weight", (co, co, *k(3))), (p + ".out_layers.3.bias", (co,)),
            ]
            if ci != co:
                out += [(p + ".skip_connection.weight", (co, ci, *k(1))), (p + ".skip_connection.bias", (co,))]
            return out
        if L["kind"] == "attn":
            c = L["ch"]
            return [
                (p + ".norm.weight", (c,)), (p + ".norm.bias", (c,)),
                (p + ".qkv.weight", (3 * c, c, 1)), (p + ".qkv.bias", (3 * c,)),
                (p + ".proj_out.weight", (c, c, 1)), (p + ".proj_out.bias", (c,)),
            ]
        raise ValueError(L["kind"])

    for blk in plan["input"]:
        for L in blk:
            specs += layer_specs(L)
    for L in plan["middle"]:
        specs += layer_specs(L)
    for blk in plan["output"]:
        for L in blk:
            specs += layer_specs(L)
    specs += [
        ("out.0.weight", (plan["out_norm_ch"],)), ("out.0.bias", (plan["out_norm_ch"],)),
        ("out.2.weight", (cfg.out_channels, plan["out_conv_in"], *k(3))), ("out.2.bias", (cfg.out_channels,)),
    ]
    return specs


# ----------------------------------------------------------------------------
# Functional forward
# ----------------------------------------------------------------------------

def timestep_embedding(t: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """guided_diffusion/nn.py:103-121 -- cos block first, then sin."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    ang = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(x, sd, p):
    """nn.py:17-19 GroupNorm32(32, C): fp32 statistics, eps=1e-5."""
    return F.group_norm(x.float(), 32, sd[p + ".weight"], sd[p + ".bias"], 1e-5).type(x.dtype)


def _conv3(x, sd, p, stride=1):
    """conv_nd (nn.py:22-32): Conv3d for (B,C,Z,H,W), Conv2d for (B,C,H,W); `stride` applies to (H, W)."""
    if x.dim() == 4:
        return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=1)
    st = (1, stride, stride) if isinstance(stride, int) and stride != 1 else stride
    return F.conv3d(x, sd[p + ".weight"], sd[p + ".bias"], stride=st, padding=1)


def _up_hw(x):
    """unet.py:100-108 nearest x2: on (H, W) only for dims=3, scale_factor=2 for dims=2."""
    if x.dim() == 4:
        return F.interpolate(x, scale_factor=2, mode="nearest")
    return F.interpolate(x, (x.shape[2], x.shape[3] * 2, x.shape[4] * 2), mode="nearest")


def _down_hw(x):
    """unet.py:129,136-137 AvgPool3d((1,2,2)) / AvgPool2d(2)."""
    if x.dim() == 4:
        return F.avg_pool2d(x, kernel_size=2, stride=2)
    return F.avg_pool3d(x, kernel_size=(1, 2, 2), stride=(1, 2, 2))


def _resblock(x, emb, sd, L, cfg):
    """unet.py:236-256 ResBlock._forward."""
    p = L["prefix"]
    h = F.silu(_gn(x, sd, p + ".in_layers.0"))
    if L["up"]:
        h, x = _up_hw(h), _up_hw(x)
    elif L["down"]:
        h, x = _down_hw(h), _down_hw(x)
    h = _conv3(h, sd, p + ".in_layers.2")
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"]).type(h.dtype)
    e = e[(..., *([None] * (h.dim() - 2)))]
    if cfg.use_scale_shift_norm:
        scale, shift = torch.chunk(e, 2, dim=1)
        h = _gn(h, sd, p + ".out_layers.0") * (1 + scale) + shift
        h = F.silu(h)
    else:
        h = F.silu(_gn(h + e, sd, p + ".out_layers.0"))
    h = _conv3(h, sd, p + ".out_layers.3")
    if L["cin"] != L["cout"]:
        skip = F.conv2d if x.dim() == 4 else F.conv3d
        x = skip(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def _attention(x, sd, L, cfg):
    """unet.py:296-305 AttentionBlock._forward with QKVAttentionLegacy
    (:328-358) or QKVAttention (:361-393)."""
    p = L["prefix"]
    b, c = x.shape[:2]
    spatial = x.shape[2:]
    xf = x.reshape(b, c, -1)
    n = xf.shape[-1]
    qkv = F.conv1d(_gn(xf, sd, p + ".norm"), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    nh = L["heads"]
    ch = c // nh
    s = 1 / math.sqrt(math.sqrt(ch))
    if cfg.use_new_attention_order:
        q, k, v = qkv.chunk(3, dim=1)
        q = (q * s).reshape(b * nh, ch, n)
        k = (k * s).reshape(b * nh, ch, n)
        v = v.reshape(b * nh, ch, n)
    else:
        q, k, v = qkv.reshape(b * nh, ch * 3, n).split(ch, dim=1)
        q, k = q * s, k * s
    w = torch.einsum("bct,bcs->bts", q, k)
    w = torch.softmax(w.float(), dim=-1).type(w.dtype)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, n)
    h = F.conv1d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + h).reshape(b, c, *spatial)


def _run_layers(h, emb, sd, layers, cfg):
    for L in layers:
        k = L["kind"]
        if k == "res":
            h = _resblock(h, emb, sd, L, cfg)
        elif k == "attn":
            h = _attention(h, sd, L, cfg)
        elif k == "conv":
            h = _conv3(h, sd, L["prefix"], stride=L["stride_hw"])
        elif k == "upconv":
            h = _conv3(_up_hw(h), sd, L["prefix"])
        else:
            raise ValueError(k)
    return h


@torch.no_grad()
def unet_forward(cfg: UNetConfig, sd: dict, x: torch.Tensor, t: torch.Tensor,
                 low_res: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None,
                 taps: Optional[dict] = None, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """SuperResModel_noatt.forward (unet.py:1687-1694) -> UNetModel_noatt.forward
    (:1015-1044).  x, low_res: (B,1,Z,H,W) fp32; t: (B,) already mapped to the
    ORIGINAL timestep numbering; returns (B,out_channels,Z,H,W) fp32.
    `taps`, if given, receives named intermediate tensors (for kernel tests).
    `dtype` (torch.float16 / bfloat16): the reference's use_fp16 flow -- h.type(self.dtype) before the input blocks
    (unet.py:1035) and h.type(x.dtype) before `out` (:1043); `sd` must then hold the torso's conv tensors in that
    dtype (fp16_util.py:15-22).  Half-precision pooling only runs on CUDA, so this is a GPU-side cross-check."""
    plan = build_plan(cfg)
    assert x.dim() == cfg.dims + 2 and (low_res is not None) == cfg.concat_low_res
    h = torch.cat([x, low_res], dim=1) if cfg.concat_low_res else x
    if dtype is not None:
        h = h.type(dtype)
    emb = timestep_embedding(t, cfg.model_channels)
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    if cfg.num_classes is not None:
        assert y is not None and y.shape == (x.shape[0],)
        emb = emb + sd["label_emb.weight"][y]
    if taps is not None:
        taps["emb"] = emb
    skips = []
    for i, blk in enumerate(plan["input"]):
        h = _run_layers(h, emb, sd, blk, cfg)
        skips.append(h)
        if taps is not None:
            taps[f"input_blocks.{i}"] = h
    h = _run_layers(h, emb, sd, plan["middle"], cfg)
    if taps is not None:
        taps["middle_block"] = h
    for i, blk in enumerate(plan["output"]):
        h = torch.cat([h, skips.pop()], dim=1)
        h = _run_layers(h, emb, sd, blk, cfg)
        if taps is not None:
            taps[f"output_blocks.{i}"] = h
    h = h.type(x.dtype)
    h = F.silu(_gn(h, sd, "out.0"))
    return _conv3(h, sd, "out.2")
