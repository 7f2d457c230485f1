"""Oracle VAE decoder: PyTorch restatement of diffusers ``AutoencoderKL.decode`` (SD-v1 config).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker of the native decoder
(sonicdiffusionbayeslab_b200/vae_engine.py) and the stock-PyTorch (cuDNN) timing baseline.  The
reference calls ``self.vae.decode`` at /root/reference/src/models.py:288-302; the module lives in
diffusers 0.32.1 (absent), so the published architecture is restated with
diffusers' state-dict key names -- the same names ``sonicdiffusionbayeslab_b200/vae_spec.py`` lists,
so one state dict feeds both.  UNPINNED (third-party layer).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Res(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(32, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class _Attn(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, c, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Identity()])

    def forward(self, x):
        B, Cc, H, W = x.shape
        h = self.group_norm(x).view(B, Cc, H * W).transpose(1, 2)
        q, k, v = self.to_q(h)[:, None], self.to_k(h)[:, None], self.to_v(h)[:, None]
        o = F.scaled_dot_product_attention(q, k, v)[:, 0]
        return x + self.to_out[0](o).transpose(1, 2).reshape(B, Cc, H, W)


class _Mid(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.resnets = nn.ModuleList([_Res(c, c), _Res(c, c)])
        self.attentions = nn.ModuleList([_Attn(c)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _Up(nn.Module):
    def __init__(self, cin, cout, up):
        super().__init__()
        self.resnets = nn.ModuleList([_Res(cin if i == 0 else cout, cout) for i in range(3)])
        self.upsamplers = None
        if up:
            conv = nn.Module()
            conv.conv = nn.Conv2d(cout, cout, 3, padding=1)
            self.upsamplers = nn.ModuleList([conv])

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0].conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class _Decoder(nn.Module):
    def __init__(self, latent=4, out=3, chans=(128, 256, 512, 512)):
        super().__init__()
        rev = list(reversed(chans))
        self.conv_in = nn.Conv2d(latent, rev[0], 3, padding=1)
        self.mid_block = _Mid(rev[0])
        ups, c = [], rev[0]
        for i, co in enumerate(rev):
            ups.append(_Up(c, co, up=i != len(rev) - 1))
            c = co
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(32, chans[0], eps=1e-6)
        self.conv_out = nn.Conv2d(chans[0], out, 3, padding=1)

    def forward(self, z):
        h = self.mid_block(self.conv_in(z))
        for u in self.up_blocks:
            h = u(h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class VaeConfig(dict):
    __getattr__ = dict.__getitem__


class AutoencoderKLDecoder(nn.Module):
    """``decode(z)`` of diffusers' AutoencoderKL (SD-v1 config, scaling_factor 0.18215)."""

    def __init__(self, chans=(128, 256, 512, 512)):
        super().__init__()
        self.post_quant_conv = nn.Conv2d(4, 4, 1)
        self.decoder = _Decoder(chans=chans)
        self.config = VaeConfig(scaling_factor=0.18215, latent_channels=4, block_out_channels=tuple(chans))

    @torch.no_grad()
    def decode(self, z, return_dict=False, generator=None):
        p = next(self.parameters())
        img = self.decoder(self.post_quant_conv(z.to(p.dtype)))
        return (img,)

    def load_diffusers_state_dict(self, sd):
        own = self.state_dict()
        missing = [k for k in own if k not in sd]
        if missing:
            raise KeyError(f"VAE checkpoint misses decoder keys, e.g. {missing[:3]}")
        self.load_state_dict({k: sd[k] for k in own})


def make_vae(seed: int = 29, dtype=torch.bfloat16, device="cpu", chans=(128, 256, 512, 512)):
    state = torch.random.get_rng_state()
    torch.manual_seed(seed + 1)
    try:
        vae = AutoencoderKLDecoder(chans)
    finally:
        torch.random.set_rng_state(state)
    return vae.to(device=device, dtype=dtype).eval()
