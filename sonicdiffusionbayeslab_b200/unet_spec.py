"""Parameter inventory of SD-v1.x ``UNet2DConditionModel`` in diffusers' state-dict layout.

Used to (a) validate a provided checkpoint before it is packed for the engine and (b) create
random-init weights when no checkpoint can be had (there is no network in this deployment;
BASELINE.json allows "random-init or provided weights").  Key names: SURVEY.md appendix A.7.
859,520,964 parameters for the default architecture.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from .unet_engine import UNetArch


def unet_param_shapes(a: UNetArch = UNetArch()) -> "OrderedDict[str, tuple]":
    s = OrderedDict()
    boc = a.block_out_channels
    T = 4 * boc[0]

    def conv(p, co, ci, k):
        s[p + ".weight"] = (co, ci, k, k)
        s[p + ".bias"] = (co,)

    def lin(p, co, ci, bias=True):
        s[p + ".weight"] = (co, ci)
        if bias:
            s[p + ".bias"] = (co,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def resnet(p, ci, co):
        norm(p + ".norm1", ci)
        conv(p + ".conv1", co, ci, 3)
        lin(p + ".time_emb_proj", co, T)
        norm(p + ".norm2", co)
        conv(p + ".conv2", co, co, 3)
        if ci != co:
            conv(p + ".conv_shortcut", co, ci, 1)

    def transformer(p, c):
        norm(p + ".norm", c)
        conv(p + ".proj_in", c, c, 1)
        t = p + ".transformer_blocks.0"
        norm(t + ".norm1", c)
        for n in ("to_q", "to_k", "to_v"):
            lin(f"{t}.attn1.{n}", c, c, bias=False)
        lin(t + ".attn1.to_out.0", c, c)
        norm(t + ".norm2", c)
        lin(t + ".attn2.to_q", c, c, bias=False)
        lin(t + ".attn2.to_k", c, a.cross_attention_dim, bias=False)
        lin(t + ".attn2.to_v", c, a.cross_attention_dim, bias=False)
        lin(t + ".attn2.to_out.0", c, c)
        norm(t + ".norm3", c)
        lin(t + ".ff.net.0.proj", 8 * c, c)
        lin(t + ".ff.net.2", c, 4 * c)
        conv(p + ".proj_out", c, c, 1)

    conv("conv_in", boc[0], a.in_channels, 3)
    lin("time_embedding.linear_1", T, boc[0])
    lin("time_embedding.linear_2", T, T)
    c = boc[0]
    for b, co in enumerate(boc):
        for j in range(a.layers_per_block):
            resnet(f"down_blocks.{b}.resnets.{j}", c if j == 0 else co, co)
            if a.attn_blocks[b]:
                transformer(f"down_blocks.{b}.attentions.{j}", co)
        if b != len(boc) - 1:
            conv(f"down_blocks.{b}.downsamplers.0.conv", co, co, 3)
        c = co
    resnet("mid_block.resnets.0", boc[-1], boc[-1])
    transformer("mid_block.attentions.0", boc[-1])
    resnet("mid_block.resnets.1", boc[-1], boc[-1])
    rev = list(reversed(boc))
    rev_attn = list(reversed(a.attn_blocks))
    c = rev[0]
    for b, co in enumerate(rev):
        skip_last = rev[min(b + 1, len(boc) - 1)]
        n = a.layers_per_block + 1
        for j in range(n):
            skip = skip_last if j == n - 1 else co
            resnet(f"up_blocks.{b}.resnets.{j}", (c if j == 0 else co) + skip, co)
            if rev_attn[b]:
                transformer(f"up_blocks.{b}.attentions.{j}", co)
        if b != len(boc) - 1:
            conv(f"up_blocks.{b}.upsamplers.0.conv", co, co, 3)
        c = co
    norm("conv_norm_out", boc[0])
    conv("conv_out", a.out_channels, boc[0], 3)
    return s


def validate_state_dict(sd, a: UNetArch = UNetArch()):
    spec = unet_param_shapes(a)
    missing = [k for k in spec if k not in sd]
    if missing:
        raise KeyError(f"UNet checkpoint misses {len(missing)} tensors, e.g. {missing[:3]}")
    bad = [(k, tuple(sd[k].shape), v) for k, v in spec.items() if tuple(sd[k].shape) != v]
    if bad:
        raise ValueError(f"UNet checkpoint shape mismatch, e.g. {bad[:3]}")


def random_unet_state_dict(seed: int = 29, a: UNetArch = UNetArch(), device="cpu"):
    """Seeded random weights: N(0, 1/fan_in) matrices (residual-branch outputs halved), unit norms."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = OrderedDict()
    for k, shape in unet_param_shapes(a).items():
        is_norm = ".norm" in k or k.startswith("conv_norm_out")
        if k.endswith(".weight") and not is_norm:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            w = torch.randn(shape, generator=g) * (fan_in ** -0.5)
            if k.endswith(("conv2.weight", "to_out.0.weight", "ff.net.2.weight", "proj_out.weight")):
                w *= 0.5
            sd[k] = w
        elif k.endswith(".weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            sd[k] = 0.1 * torch.randn(shape, generator=g) if is_norm else 0.02 * torch.randn(shape, generator=g)
    return OrderedDict((k, v.to(device)) for k, v in sd.items())
