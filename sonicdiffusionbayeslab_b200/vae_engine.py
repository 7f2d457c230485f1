"""B200 VAE decoder engine: ``AutoencoderKL.decode`` (SD-v1 config) recorded into a native launch plan.

The reference decodes the final latents (and every x0 prediction) with ``self.vae.decode`` after the denoising loop
(/root/reference/src/models.py:288-302); SURVEY.md section 8(f) ranks it first among the "next" rows (2.51 TFLOP per
image).  This engine reuses the UNet kernels: tcgen05 implicit-GEMM 3x3 convolutions up to 512x512 (TMA boxes of 128
pixels along a row), GroupNorm(+SiLU) with statistics from the producing GEMM's epilogue, nearest-2x upsampling.
The single-head d=512 attention of the mid block (4096 tokens) is two GEMMs around an in-place row softmax
(``sonic_softmax_rows``): S_b = Q_b K_b^T, P_b = softmax(S_b / sqrt(512)), O_b = P_b V_b with V^T produced directly
as ``W_v X_b^T`` (so no transpose kernel) and the value bias added after the product (rows of P sum to one).

State-dict keys are diffusers' ``AutoencoderKL`` names (``post_quant_conv.*``, ``decoder.*``), i.e. what
``vae.AutoencoderKLDecoder.state_dict()`` yields and what a real SD-v1.5 ``vae`` checkpoint holds.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import kernels as K
from ._lib import check, lib
from .unet_engine import Arena, UNetEngine, _Plan


class _VaeArch:
    norm_num_groups = 32


class VaeEngine(UNetEngine):
    """One engine = one (batch, latent size, io dtype) specialisation of the decoder."""

    def __init__(self, state_dict, *, n_img: int, latent: int = 64, chans=(128, 256, 512, 512),
                 io_dtype=torch.bfloat16, device="cuda"):
        # deliberately not calling UNetEngine.__init__: only its recording helpers are reused
        self.sd = state_dict
        self.dev = torch.device(device)
        self.n = n_img
        self.arch = _VaeArch()
        self.io_dtype = io_dtype
        self.arena = Arena(self.dev)
        self.fuse_gn_stats = True
        self._w = {}
        self.latent = latent
        self.chans = tuple(chans)
        self.z_in = torch.zeros(n_img, 4, latent, latent, device=self.dev, dtype=io_dtype)
        self.img = torch.zeros(n_img, 3, 8 * latent, 8 * latent, device=self.dev, dtype=io_dtype)
        with torch.no_grad():
            self.plan = self._build()
        self._graph_stream = None

    # ------------------------------------------------------------------ blocks
    def _res(self, plan, prefix, x, H, W, cout):
        cin = x.shape[-1]
        h = self._gn(plan, x, None, prefix + ".norm1", H * W, 1e-6, True)
        h1 = self._gemm(plan, h, self._conv3(prefix + ".conv1.weight"), cout, n_img=self.n, H=H, W=W, taps=9,
                        bias=self._f32(prefix + ".conv1.bias"), gn_stats=True)
        self.arena.release(h)
        h2 = self._gn(plan, h1, None, prefix + ".norm2", H * W, 1e-6, True)
        self.arena.release(h1)
        res = x
        if cin != cout:
            res = self._gemm(plan, x, self._lin(prefix + ".conv_shortcut.weight"), cout, n_img=self.n, H=H, W=W,
                             bias=self._f32(prefix + ".conv_shortcut.bias"))
        out = self._gemm(plan, h2, self._conv3(prefix + ".conv2.weight"), cout, n_img=self.n, H=H, W=W, taps=9,
                         bias=self._f32(prefix + ".conv2.bias"), residual=res, gn_stats=True)
        self.arena.release(h2)
        if res is not x:
            self.arena.release(res)
        return out

    def _mid_attention(self, plan, prefix, x, H, W):
        """x <- x + to_out(softmax(q k^T / sqrt(C)) v), one head of width C over H*W tokens per image."""
        Cc, S = x.shape[-1], H * W
        g = self._gn(plan, x, None, prefix + ".group_norm", S, 1e-6, False)
        q = self._gemm(plan, g, self._lin(prefix + ".to_q.weight"), Cc, bias=self._f32(prefix + ".to_q.bias"))
        k = self._gemm(plan, g, self._lin(prefix + ".to_k.weight"), Cc, bias=self._f32(prefix + ".to_k.bias"))
        scores = self.arena.alloc((S, S))
        vt = self.arena.alloc((Cc, S))
        o = self.arena.alloc((self.n * S, Cc))
        wv = self._lin(prefix + ".to_v.weight")                      # [C, C]: the "activation" of the V^T product
        for b in range(self.n):
            rows = slice(b * S, (b + 1) * S)
            self._gemm(plan, q[rows], k[rows], S, out=scores)        # S_b = Q_b K_b^T          [S, S]
            check(lib().sonic_plan_add_softmax_rows(plan.h, K.ptr(scores), S, S, C.c_int64(S), C.c_float(Cc ** -0.5)),
                  "sonic_plan_add_softmax_rows")
            plan.log.append(f"softmax_rows {S}x{S}")
            self._gemm(plan, wv, g[rows], S, out=vt)                 # V_b^T = W_v X_b^T         [C, S]
            self._gemm(plan, scores, vt, Cc, bias=self._f32(prefix + ".to_v.bias"), out=o[rows])   # P_b V_b + b_v
        for t in (scores, vt, q, k, g):
            self.arena.release(t)
        self._gemm(plan, o, self._lin(prefix + ".to_out.0.weight"), Cc, bias=self._f32(prefix + ".to_out.0.bias"),
                   residual=x, out=x, gn_stats=True)
        self.arena.release(o)
        return x

    # ------------------------------------------------------------------ plan
    def _build(self):
        plan = _Plan()
        n, H, W = self.n, self.latent, self.latent
        rev = list(reversed(self.chans))
        # latents NCHW -> NHWC (8 channels, zero padded)
        z8 = self.arena.alloc((n * H * W, 8))
        check(lib().sonic_plan_add_nchw_to_nhwc8(plan.h, K.ptr(self.z_in), K._dtype_code(self.z_in), n, 4, H * W, 0,
                                                 K.ptr(z8)), "sonic_plan_add_nchw_to_nhwc8")
        plan.log.append("nchw_to_nhwc8")
        # post_quant_conv (1x1, 4 -> 4), padded to K = 8 inputs / N = 16 outputs (extra rows are zero)
        wpq = torch.zeros(16, 8, device=self.dev, dtype=torch.bfloat16)
        wpq[:4, :4] = self._p("post_quant_conv.weight").reshape(4, 4).to(torch.bfloat16)
        bpq = torch.zeros(16, device=self.dev, dtype=torch.float32)
        bpq[:4] = self._p("post_quant_conv.bias").float()
        self._w["pq"] = (wpq, bpq)
        zq = self._gemm(plan, z8, wpq, 16, bias=bpq, block_n=16)
        self.arena.release(z8)
        # conv_in reads the first 8 of the 16 channels
        w = self._p("decoder.conv_in.weight")
        wp = torch.zeros(w.shape[0], 8, 3, 3, device=self.dev, dtype=w.dtype)
        wp[:, :4] = w
        self._w["conv_in"] = K.pack_conv3x3_weight(wp)
        h = self._gemm(plan, zq, self._w["conv_in"], rev[0], n_img=n, H=H, W=W, taps=9, c0=8,
                       bias=self._f32("decoder.conv_in.bias"), gn_stats=True)
        self.arena.release(zq)
        # mid block
        r = self._res(plan, "decoder.mid_block.resnets.0", h, H, W, rev[0])
        self.arena.release(h)
        r = self._mid_attention(plan, "decoder.mid_block.attentions.0", r, H, W)
        h = self._res(plan, "decoder.mid_block.resnets.1", r, H, W, rev[0])
        self.arena.release(r)
        # up blocks
        for i, cout in enumerate(rev):
            for j in range(3):
                r = self._res(plan, f"decoder.up_blocks.{i}.resnets.{j}", h, H, W, cout)
                self.arena.release(h)
                h = r
            if i != len(rev) - 1:
                src = h                       # nearest 2x + 3x3 as four phase-wise 2x2 convolutions of the source
                h, H, W = self._upsample_conv(plan, f"decoder.up_blocks.{i}.upsamplers.0", src, H, W, cout)
                self.arena.release(src)
        # out: GroupNorm + SiLU -> conv3x3 (3 output channels padded to one 16-wide MMA tile) -> NCHW
        g = self._gn(plan, h, None, "decoder.conv_norm_out", H * W, 1e-6, True)
        self.arena.release(h)
        w = self._p("decoder.conv_out.weight")
        wp = torch.zeros(16, w.shape[1], 3, 3, device=self.dev, dtype=w.dtype)
        wp[:3] = w
        bp = torch.zeros(16, device=self.dev, dtype=torch.float32)
        bp[:3] = self._p("decoder.conv_out.bias").float()
        self._w["conv_out"] = (K.pack_conv3x3_weight(wp), bp)
        o16 = self._gemm(plan, g, self._w["conv_out"][0], 16, n_img=n, H=H, W=W, taps=9, bias=bp, block_n=16)
        self.arena.release(g)
        check(lib().sonic_plan_add_nhwc_to_nchw(plan.h, K.ptr(o16), 16, n, 3, H * W, K.ptr(self.img),
                                                K._dtype_code(self.img)), "sonic_plan_add_nhwc_to_nchw")
        plan.log.append("nhwc_to_nchw")
        self.arena.release(o16)
        return plan

    # ------------------------------------------------------------------ execution
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """z: (n_img, 4, latent, latent), already divided by the scaling factor; returns (n_img, 3, 8L, 8L)."""
        assert z.shape == self.z_in.shape, (z.shape, self.z_in.shape)
        self.z_in.copy_(z)
        self.plan.run(K.stream_ptr())
        return self.img

    def stats(self, name="full"):
        return self.plan.stats()
