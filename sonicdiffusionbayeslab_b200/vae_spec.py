"""VAE decoder weights: tensor names / shapes of diffusers' ``AutoencoderKL`` decoder (SD-v1 config) and a
seeded random initialiser.

The decoder itself is the native launch plan of ``vae_engine.VaeEngine`` (call sites
/root/reference/src/models.py:288-302); this module only describes what it loads, so that a real SD-v1.5
``vae/diffusion_pytorch_model.safetensors`` is validated before it reaches the kernels and a run without weights
gets reproducible random ones ("random-init or provided weights").
Architecture: SURVEY.md appendix A.8.
"""
from __future__ import annotations

from collections import OrderedDict
from types import SimpleNamespace

import torch

SCALING_FACTOR = 0.18215


def vae_decoder_param_shapes(chans=(128, 256, 512, 512), latent=4, out=3) -> "OrderedDict[str, tuple]":
    s = OrderedDict()

    def conv(p, co, ci, k):
        s[p + ".weight"], s[p + ".bias"] = (co, ci, k, k), (co,)

    def lin(p, co, ci):
        s[p + ".weight"], s[p + ".bias"] = (co, ci), (co,)

    def norm(p, c):
        s[p + ".weight"], s[p + ".bias"] = (c,), (c,)

    def res(p, ci, co):
        norm(p + ".norm1", ci)
        conv(p + ".conv1", co, ci, 3)
        norm(p + ".norm2", co)
        conv(p + ".conv2", co, co, 3)
        if ci != co:
            conv(p + ".conv_shortcut", co, ci, 1)

    rev = list(reversed(chans))
    conv("post_quant_conv", latent, latent, 1)
    conv("decoder.conv_in", rev[0], latent, 3)
    res("decoder.mid_block.resnets.0", rev[0], rev[0])
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", rev[0])
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        lin(f"{a}.{n}", rev[0], rev[0])
    res("decoder.mid_block.resnets.1", rev[0], rev[0])
    c = rev[0]
    for i, co in enumerate(rev):
        for j in range(3):
            res(f"decoder.up_blocks.{i}.resnets.{j}", c if j == 0 else co, co)
        if i != len(rev) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", co, co, 3)
        c = co
    norm("decoder.conv_norm_out", chans[0])
    conv("decoder.conv_out", out, chans[0], 3)
    return s


def validate_vae_state_dict(sd, chans=(128, 256, 512, 512)):
    """Returns the decoder subset of ``sd`` (an AutoencoderKL checkpoint also holds the encoder)."""
    spec = vae_decoder_param_shapes(chans)
    missing = [k for k in spec if k not in sd]
    if missing:
        raise KeyError(f"VAE checkpoint misses decoder keys, e.g. {missing[:3]}")
    out = OrderedDict()
    for k, shape in spec.items():
        t = sd[k]
        if t.dim() == 4 and len(shape) == 2:          # older checkpoints store the attention projections as 1x1 convs
            t = t.reshape(t.shape[0], t.shape[1])
        if tuple(t.shape) != shape:
            raise ValueError(f"VAE checkpoint shape mismatch at {k}: {tuple(t.shape)} != {shape}")
        out[k] = t
    return out


def random_vae_state_dict(seed: int = 29, chans=(128, 256, 512, 512)):
    """Seeded weights with PyTorch's default scale (uniform +-1/sqrt(fan_in)), unit norms."""
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    sd = OrderedDict()
    for k, shape in vae_decoder_param_shapes(chans).items():
        is_norm = "norm" in k.rsplit(".", 2)[-2]
        if is_norm:
            sd[k] = torch.ones(shape) if k.endswith(".weight") else torch.zeros(shape)
            continue
        base = k.rsplit(".", 1)[0] + ".weight"
        fan_in = 1
        for d in vae_decoder_param_shapes(chans)[base][1:]:
            fan_in *= d
        bound = fan_in ** -0.5
        sd[k] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


class VaeWeights:
    """What the pipelines keep in ``pipe.vae``: the decoder state dict and ``config.scaling_factor``."""

    def __init__(self, state_dict, chans=(128, 256, 512, 512)):
        self._sd = state_dict
        self.config = SimpleNamespace(scaling_factor=SCALING_FACTOR, latent_channels=4, block_out_channels=tuple(chans))

    def state_dict(self):
        return self._sd

    def to(self, *_a, **_k):
        return self
