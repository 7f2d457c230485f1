"""Oracle UNet: PyTorch restatement of diffusers ``UNet2DConditionModel`` (SD-v1.x config).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference calls this network at
/root/reference/src/models.py:227-235; its source lives in diffusers 0.32.1
(poetry.lock:454-455), which is absent here, so the published architecture is restated
from SURVEY.md appendix A.1 with diffusers' state-dict key names (appendix A.7) so a
diffusers-layout checkpoint loads unchanged.  Parameter count of the default config is
859,520,964 -- the public SD-v1.5 figure -- which is the only available pin.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    sample_size: int = 64
    block_out_channels: tuple = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    attn_blocks: tuple = (True, True, True, False)      # CrossAttnDownBlock2D x3, DownBlock2D
    num_heads: int = 8                                   # config.attention_head_dim=8 means 8 heads
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    time_cond_proj_dim: object = None


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0), float32."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout) if temb_dim else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, dim, heads, ctx_dim=None):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim or dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim or dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        B, S, C = x.shape
        d = C // self.heads
        q = self.to_q(x).view(B, S, self.heads, d).transpose(1, 2)
        k = self.to_k(ctx).view(B, -1, self.heads, d).transpose(1, 2)
        v = self.to_v(ctx).view(B, -1, self.heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)               # scale d^-1/2 (AttnProcessor2_0)
        return self.to_out[0](o.transpose(1, 2).reshape(B, S, C))


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, ctx_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, ctx_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        return x + self.ff(self.norm3(x))


class Transformer2DModel(nn.Module):
    def __init__(self, dim, heads, ctx_dim, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, ctx_dim)])
        self.proj_out = nn.Conv2d(dim, dim, 1)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        r = x
        h = self.proj_in(self.norm(x)).permute(0, 2, 3, 1).reshape(B, H * W, C)
        for blk in self.transformer_blocks:
            h = blk(h, ctx)
        h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
        return self.proj_out(h) + r


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cfg, cin, cout, attn, down):
        super().__init__()
        T = cfg.block_out_channels[0] * 4
        self.resnets = nn.ModuleList([
            ResnetBlock2D(cin if i == 0 else cout, cout, T, cfg.norm_num_groups, cfg.norm_eps)
            for i in range(cfg.layers_per_block)])
        if attn:
            self.attentions = nn.ModuleList([
                Transformer2DModel(cout, cfg.num_heads, cfg.cross_attention_dim, cfg.norm_num_groups)
                for _ in range(cfg.layers_per_block)])
        self.has_attn = attn
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if down else None


class UpBlock(nn.Module):
    def __init__(self, cfg, cprev, cout, cin_skip_last, attn, up):
        super().__init__()
        T = cfg.block_out_channels[0] * 4
        n = cfg.layers_per_block + 1
        res = []
        for i in range(n):
            skip = cin_skip_last if i == n - 1 else cout
            res.append(ResnetBlock2D((cprev if i == 0 else cout) + skip, cout, T,
                                     cfg.norm_num_groups, cfg.norm_eps))
        self.resnets = nn.ModuleList(res)
        if attn:
            self.attentions = nn.ModuleList([
                Transformer2DModel(cout, cfg.num_heads, cfg.cross_attention_dim, cfg.norm_num_groups)
                for _ in range(n)])
        self.has_attn = attn
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if up else None


class MidBlock(nn.Module):
    def __init__(self, cfg, c):
        super().__init__()
        T = cfg.block_out_channels[0] * 4
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, T, cfg.norm_num_groups, cfg.norm_eps) for _ in range(2)])
        self.attentions = nn.ModuleList([
            Transformer2DModel(c, cfg.num_heads, cfg.cross_attention_dim, cfg.norm_num_groups)])


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: UNetConfig = UNetConfig()):
        super().__init__()
        self.config = cfg
        boc = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], boc[0] * 4)
        downs, c = [], boc[0]
        for i, co in enumerate(boc):
            downs.append(DownBlock(cfg, c, co, cfg.attn_blocks[i], down=i != len(boc) - 1))
            c = co
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlock(cfg, boc[-1])
        rev = list(reversed(boc))
        rev_attn = list(reversed(cfg.attn_blocks))
        ups, c = [], rev[0]
        for i, co in enumerate(rev):
            skip_last = rev[min(i + 1, len(boc) - 1)]
            ups.append(UpBlock(cfg, c, co, skip_last, rev_attn[i], up=i != len(boc) - 1))
            c = co
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, boc[0], eps=cfg.norm_eps)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    # The forward is split into named stages so the DeepCache oracle can skip/replay them.
    def time_embed(self, sample, timestep):
        t = torch.as_tensor(timestep, device=sample.device)
        if t.dim() == 0:
            t = t[None]
        t = t.expand(sample.shape[0])
        temb = timestep_embedding(t, self.config.block_out_channels[0]).to(sample.dtype)
        return self.time_embedding(temb)

    def run_down_layer(self, blk, j, h, temb, ctx):
        h = blk.resnets[j](h, temb)
        if blk.has_attn:
            h = blk.attentions[j](h, ctx)
        return h

    def run_up_layer(self, blk, j, h, skip, temb, ctx):
        h = blk.resnets[j](torch.cat([h, skip], dim=1), temb)
        if blk.has_attn:
            h = blk.attentions[j](h, ctx)
        return h

    def forward(self, sample, timestep, encoder_hidden_states, return_dict=False, **_):
        temb = self.time_embed(sample, timestep)
        ctx = encoder_hidden_states
        h = self.conv_in(sample)
        skips = [h]
        for blk in self.down_blocks:
            for j in range(len(blk.resnets)):
                h = self.run_down_layer(blk, j, h, temb, ctx)
                skips.append(h)
            if blk.downsamplers is not None:
                h = blk.downsamplers[0](h)
                skips.append(h)
        h = self.mid_block.resnets[0](h, temb)
        h = self.mid_block.attentions[0](h, ctx)
        h = self.mid_block.resnets[1](h, temb)
        for blk in self.up_blocks:
            for j in range(len(blk.resnets)):
                h = self.run_up_layer(blk, j, h, skips.pop(), temb, ctx)
            if blk.upsamplers is not None:
                h = blk.upsamplers[0](h)
        h = self.conv_out(F.silu(self.conv_norm_out(h)))
        return (h,)


def make_unet(seed: int = 29, cfg: UNetConfig = UNetConfig(), dtype=torch.float32) -> UNet2DConditionModel:
    """Random-init weights under ``torch.manual_seed(seed)``.

    PyTorch default inits leave the residual stream growing with depth, which is fine for
    fp32 but makes bf16 comparisons needlessly loose, so the output projection of every
    residual branch (conv2, to_out, ff.net.2, proj_out) is scaled down; the function is the
    single source of weights for both oracle and engine so parity is unaffected.
    """
    g = torch.Generator().manual_seed(seed)
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        net = UNet2DConditionModel(cfg)
    finally:
        torch.random.set_rng_state(state)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith(("conv2.weight", "to_out.0.weight", "ff.net.2.weight", "proj_out.weight")):
                p.mul_(0.5)
            if name.endswith("norm.weight") or ".norm1.weight" in name or ".norm2.weight" in name \
                    or ".norm3.weight" in name or name == "conv_norm_out.weight":
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            if name.endswith("norm.bias") or ".norm1.bias" in name or ".norm2.bias" in name \
                    or ".norm3.bias" in name or name == "conv_norm_out.bias":
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return net.to(dtype).eval()
