"""CPU: where the 16-bit error of the UNet comes from, measured on the oracle with storage roundings emulated at
chosen points (weights / conv-operand activations / the tensor between a ResBlock's convs / block outputs).

This is the evidence behind the storage formats of the default bf16 mode:
  * every one of the four rounding points costs about the same, so an all-bf16 network sits at 1.2 .. 2.5e-2 eps max-rel
    on random-weight networks -- above the north star's 1e-2;
  * the two tensors that are NOT operands of the dense contraction (the tensor between the two convs, block inputs /
    outputs) can be stored as fp16 at no cost: error x ~0.65;
  * rounding ONLY the convolution weights to bf16 already costs 1.0e-2 on the C1 toy network, so no design with bf16
    tensor-core operands reaches 1e-2 there -- that needs fp16 operands (the reference's dtype), which measure 2e-3.
The CUDA path's measured numbers (tests/test_gpu_model.py) track these emulated ones to ~10 %.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import cases, unet as U
from oracle.weights import synth_inputs, synth_state_dict

FMT = {"bf16": torch.bfloat16, "fp16": torch.float16}


def emulated_forward(cfg, sd, x, t, low, mode):
    """oracle.unet.unet_forward with `mode[point]` in {None, "bf16", "fp16"} applied at point in
    {"w", "w_io", "act", "mid", "trunk"}: 3x3x3 / qkv / proj weights, weights of convs that read a block input (skip,
    stem, Downsample / Upsample), conv-operand activations, the tensor between a block's convs, block outputs."""

    def rnd(v, point):
        k = mode.get(point)
        return v if k is None else v.to(FMT[k]).float()

    def resblock(xx, emb, sd, L, cfg):
        p = L["prefix"]
        h = F.silu(U._gn(xx, sd, p + ".in_layers.0"))
        if L["up"]:
            h, xx = U._up_hw(h), U._up_hw(xx)
        elif L["down"]:
            h, xx = U._down_hw(h), U._down_hw(xx)
        conv = F.conv2d if xx.dim() == 4 else F.conv3d
        h = rnd(conv(rnd(h, "act"), rnd(sd[p + ".in_layers.2.weight"], "w"), sd[p + ".in_layers.2.bias"], padding=1), "mid")
        e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
        e = e[(..., *([None] * (h.dim() - 2)))]
        if cfg.use_scale_shift_norm:
            scale, shift = torch.chunk(e, 2, dim=1)
            h = F.silu(U._gn(h, sd, p + ".out_layers.0") * (1 + scale) + shift)
        else:
            h = F.silu(U._gn(h + e, sd, p + ".out_layers.0"))
        h = conv(rnd(h, "act"), rnd(sd[p + ".out_layers.3.weight"], "w"), sd[p + ".out_layers.3.bias"], padding=1)
        if L["cin"] != L["cout"]:
            xx = conv(xx, rnd(sd[p + ".skip_connection.weight"], "w_io"), sd[p + ".skip_connection.bias"])
        return rnd(xx + h, "trunk")

    conv3_orig = U._conv3

    def conv3(xx, sd, p, stride=1):
        if p == "out.2":  # the fp32 head
            return conv3_orig(xx, sd, p, stride)
        sd2 = dict(sd)
        sd2[p + ".weight"] = rnd(sd[p + ".weight"], "w_io")
        if p == "input_blocks.0.0":
            xx = rnd(xx, "trunk")  # the packed network input is stored like a block input
        return rnd(conv3_orig(xx, sd2, p, stride), "trunk")

    def attention(xx, sd, L, cfg):
        p = L["prefix"]
        b, c = xx.shape[:2]
        spatial = xx.shape[2:]
        xf = xx.reshape(b, c, -1)
        n = xf.shape[-1]
        qkv = rnd(F.conv1d(rnd(U._gn(xf, sd, p + ".norm"), "act"), rnd(sd[p + ".qkv.weight"], "w"), sd[p + ".qkv.bias"]), "act")
        nh = L["heads"]
        ch = c // nh
        s = 1 / math.sqrt(math.sqrt(ch))
        if cfg.use_new_attention_order:
            q, k, v = qkv.chunk(3, dim=1)
            q, k, v = (q * s).reshape(b * nh, ch, n), (k * s).reshape(b * nh, ch, n), v.reshape(b * nh, ch, n)
        else:
            q, k, v = qkv.reshape(b * nh, ch * 3, n).split(ch, dim=1)
            q, k = q * s, k * s
        w = rnd(torch.softmax(torch.einsum("bct,bcs->bts", q, k).float(), dim=-1), "act")
        a = rnd(torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, n), "act")
        h = F.conv1d(a, rnd(sd[p + ".proj_out.weight"], "w"), sd[p + ".proj_out.bias"])
        return rnd((xf + h).reshape(b, c, *spatial), "trunk")

    saved = (U._resblock, U._conv3, U._attention)
    U._resblock, U._conv3, U._attention = resblock, conv3, attention
    try:
        return U.unet_forward(cfg, sd, x, t, low)
    finally:
        U._resblock, U._conv3, U._attention = saved


MODES = {
    "bf16_strict": dict(w="bf16", w_io="bf16", act="bf16", mid="bf16", trunk="bf16"),  # round 1: every tensor bf16
    "bf16": dict(w="bf16", w_io="fp16", act="bf16", mid="fp16", trunk="fp16"),          # the default mode
    "weights_only": dict(w="bf16", w_io="bf16"),
    "fp16": dict(w="fp16", w_io="fp16", act="fp16", mid="fp16", trunk="fp16"),
}


def errors(flags, shape, seed, t):
    cfg = cases.cfg_from_flags(cases.sr_flags(**flags))
    sd = synth_state_dict(cfg, seed=seed)
    low, x, _ = synth_inputs(shape, 0)
    tt = torch.tensor(t)
    ref = U.unet_forward(cfg, sd, x, tt, low)
    out = {}
    for name, mode in MODES.items():
        got = emulated_forward(cfg, sd, x, tt, low, mode)
        out[name] = float((got - ref).abs().max() / ref.abs().max())
    return out


def test_emulation_with_no_rounding_is_the_oracle():
    c = cases.UNET_CASES["attn"]
    cfg = cases.cfg_from_flags(cases.sr_flags(**c["flags"]))
    sd = synth_state_dict(cfg, seed=c.get("seed", 0))
    low, x, _ = synth_inputs(c["shape"], 0)
    t = torch.tensor(c["t"])
    assert torch.equal(emulated_forward(cfg, sd, x, t, low, {}), U.unet_forward(cfg, sd, x, t, low))


@pytest.mark.parametrize("name", ["wide", "plainconv"])
def test_fp16_storage_of_non_operand_tensors_meets_1e2(name):
    c = cases.UNET_CASES[name]
    e = errors(c["flags"], c["shape"], c.get("seed", 0), c["t"])
    print(name, {k: f"{v:.2e}" for k, v in e.items()})
    assert e["bf16_strict"] > 1e-2          # the round-1 design misses the north star's bound ...
    assert e["bf16"] < 1e-2                 # ... bf16 operands + fp16 storage meets it on these networks ...
    assert e["bf16"] < 0.85 * e["bf16_strict"]
    assert e["fp16"] < 3e-3                 # ... and the reference's own dtype has 4-5x margin


def test_bf16_weights_alone_exceed_1e2_on_c1():
    """BASELINE configs[0] (32 channels, 32^3): why `c1_first_eps` is listed in tests/test_gpu_model.py."""
    e = errors(cases.C1_FLAGS, cases.C1_SHAPE, 0, [999])
    print("c1", {k: f"{v:.2e}" for k, v in e.items()})
    assert e["weights_only"] > 0.9e-2      # (1.01e-2 here) the weights alone use up the whole budget ...
    assert e["bf16"] > 1.3e-2              # ... so with bf16 operand activations on top the default mode cannot meet 1e-2
    assert e["bf16"] < e["bf16_strict"]
    assert e["fp16"] < 5e-3
