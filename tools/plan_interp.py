"""A CPU interpreter for recorded launch plans: executes the operator list of a plan with plain PyTorch on the HOST
buffers the engine was recorded over (``device="cpu"``, tools/plan_check.py ``recording``), following the operator
semantics documented in include/sonic.h.  It checks what no kernel test can see: the WIRING of a plan -- weight
packing (tap-major 3x3, phase-form upsample, tile-interleaved GEGLU, folded LayerNorm side chunks), concat order, skip
connections, the time path, the DeepCache cut of every branch -- against the oracle UNet, for any batch size, without a
GPU.  Arithmetic: bf16 operands as stored, fp32 accumulation, outputs rounded to bf16 where the kernels do; it is a
model of the dataflow, not of the kernels' rounding order (tests compare with tolerances, not bit-exactly).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BF16, F32 = torch.bfloat16, torch.float32


def _mem(ptr, count, dtype):
    """A tensor view of ``count`` elements of ``dtype`` at host address ``ptr`` (shares memory: writes go through)."""
    size = count * torch.empty((), dtype=dtype).element_size()
    return torch.frombuffer((C.c_char * size).from_address(ptr), dtype=torch.uint8).view(dtype)


def _rows(ptr, rows, cols, ld, dtype=BF16):
    """[rows, cols] view with row pitch ``ld`` elements."""
    flat = _mem(ptr, (rows - 1) * ld + cols, dtype)
    return torch.as_strided(flat, (rows, cols), (ld, 1))


def _val(a):
    return a.value if hasattr(a, "value") else a


class PlanInterpreter:
    def __init__(self):
        self.ops = {}                                           # plan handle -> [(name, args)]

    def record(self, name, args):
        self.ops.setdefault(args[0].value, []).append((name, args))

    def run(self, handle):
        with torch.no_grad():
            for name, args in self.ops[_val(handle)]:
                getattr(self, name[len("sonic_plan_add_"):])(*args[1:])

    # ------------------------------------------------------------------------------------------------ GEMM / conv
    def conv_gemm(self, ref):
        g = ref._obj
        n_img, H, W, K, N = g.n_img, g.H, g.W, g.c0 + g.c1, g.N
        m_src = n_img * H * W
        rows_in = m_src * (4 if g.stride == 2 else 1)
        a = _rows(g.a0, rows_in, g.c0, g.ld0).float()
        if g.a1:
            a = torch.cat([a, _rows(g.a1, rows_in, g.c1, g.ld1).float()], dim=1)
        w = _mem(g.w, (16 if g.upsample else g.taps) * N * K, BF16).float()
        if g.taps == 1:
            acc = a @ w.view(N, K).t()
        elif g.upsample:                                        # nearest-2x + 3x3 == four 2x2 kernels on the source
            src = F.pad(a.view(n_img, H, W, K).permute(0, 3, 1, 2), (1, 1, 1, 1))
            wp = w.view(4, 4, N, K)
            out = torch.empty(n_img, N, 2 * H, 2 * W)
            for pa in (0, 1):
                for pb in (0, 1):
                    kern = wp[2 * pa + pb].view(2, 2, N, K).permute(2, 3, 0, 1)         # [N, K, ty, tx]
                    out[:, :, pa::2, pb::2] = F.conv2d(src[:, :, pa:pa + H + 1, pb:pb + W + 1], kern)
            acc = out.permute(0, 2, 3, 1).reshape(4 * m_src, N)
        else:                                                   # 3x3, pad 1, stride 1 or 2: w = [tap = 3 ky + kx][N][K]
            s = 2 if g.stride == 2 else 1
            x = a.view(n_img, H * s, W * s, K).permute(0, 3, 1, 2)
            kern = w.view(3, 3, N, K).permute(2, 3, 0, 1)
            acc = F.conv2d(x, kern, padding=1, stride=s).permute(0, 2, 3, 1).reshape(m_src, N)
        rows_out = acc.shape[0]
        if g.row_scale:                                         # folded LayerNorm: rstd per row
            acc = acc * _mem(g.row_scale, rows_out, F32)[:, None]
        if g.bias:
            acc = acc + _mem(g.bias, N, F32)[None, :]
        if g.row_bias:                                          # per image (the resnets' time embedding)
            rb = _mem(g.row_bias, n_img * N, F32).view(n_img, N)
            acc = acc + rb.repeat_interleave(rows_out // n_img, dim=0)
        if g.epilogue == 1:                                     # GEGLU, value / gate halves interleaved per block_n tile
            bn = g.block_n
            assert bn and N % bn == 0, (N, bn)
            t = acc.view(rows_out, N // bn, 2, bn // 2)
            acc = (t[:, :, 0] * F.gelu(t[:, :, 1])).reshape(rows_out, N // 2)
        elif g.epilogue == 2:                                   # QuickGELU
            acc = acc * torch.sigmoid(1.702 * acc)
        width = acc.shape[1]
        if g.residual:
            acc = acc + _rows(g.residual, rows_out, width, g.ld_res).float()
        y = acc.to(BF16)
        _rows(g.out, rows_out, width, g.ld_out).copy_(y)
        yf = y.float()
        if g.gn_partial:                                        # [ceil(M / 32)][width][2]: sums over 32-row blocks
            blocks = -(-rows_out // 32)
            pad = F.pad(yf, (0, 0, 0, blocks * 32 - rows_out)).view(blocks, 32, width)
            _mem(g.gn_partial, blocks * width * 2, F32).view(blocks, width, 2).copy_(
                torch.stack([pad.sum(1), (pad * pad).sum(1)], dim=-1))
        if g.ln_stats_out:                                      # [M][parts][2]: everything in slot 0 (ln_side adds them)
            parts = 2 * -(-N // g.block_n)
            st = _mem(g.ln_stats_out, rows_out * parts * 2, F32).view(rows_out, parts, 2)
            st.zero_()
            st[:, 0, 0], st[:, 0, 1] = yf.sum(1), (yf * yf).sum(1)

    def ln_side(self, partials, parts, m, k, eps, side, rstd):
        parts, m, k, eps = _val(parts), _val(m), _val(k), _val(eps)
        st = _mem(_val(partials), m * parts * 2, F32).view(m, parts, 2).sum(1)
        mean = st[:, 0] / k
        var = (st[:, 1] / k - mean * mean).clamp_min(0)
        r = torch.rsqrt(var + eps)
        sd = (var + eps) * r

        def hi_lo(x):
            hi = x.to(BF16)
            return hi, (x - hi.float()).to(BF16)

        (m_hi, m_lo), (d_hi, d_lo) = hi_lo(-mean), hi_lo(sd)
        s = _mem(_val(side), m * 64, BF16).view(m, 64)
        for col, v in enumerate((m_hi, m_hi, m_lo, m_lo, d_hi, d_hi, d_lo, d_lo)):
            s[:, col] = v
        _mem(_val(rstd), m, F32).copy_(r)

    # ------------------------------------------------------------------------------------------------ attention
    def attention(self, ref):
        a = ref._obj
        hd = a.heads * a.head_dim

        def heads(ptr, seq, ld):
            return _rows(ptr, a.batch * seq, hd, ld).float().view(a.batch, seq, a.heads, a.head_dim).permute(0, 2, 1, 3)

        q, k, v = heads(a.q, a.seq_q, a.ld_q), heads(a.k, a.seq_k, a.ld_k), heads(a.v, a.seq_k, a.ld_v)
        s = (q @ k.transpose(-1, -2)) * a.scale
        if a.causal:
            s = s.masked_fill(torch.ones(a.seq_q, a.seq_k, dtype=torch.bool).triu(1), float("-inf"))
        o = (torch.softmax(s, dim=-1) @ v).permute(0, 2, 1, 3).reshape(a.batch * a.seq_q, hd)
        _rows(a.o, a.batch * a.seq_q, hd, a.ld_o).copy_(o.to(BF16))

    # ------------------------------------------------------------------------------------------------ norms
    def _groupnorm(self, x0, c0, x1, c1, n_img, hw, groups, eps, gamma, beta, silu, y):
        c = c0 + (c1 if x1 else 0)
        x = _mem(x0, n_img * hw * c0, BF16).view(n_img, hw, c0).float()
        if x1:
            x = torch.cat([x, _mem(x1, n_img * hw * c1, BF16).view(n_img, hw, c1).float()], dim=2)
        out = F.group_norm(x.permute(0, 2, 1), groups, _mem(gamma, c, F32), _mem(beta, c, F32), eps)
        if silu:
            out = F.silu(out)
        _mem(y, n_img * hw * c, BF16).view(n_img, hw, c).copy_(out.permute(0, 2, 1).to(BF16))

    def groupnorm(self, x0, c0, x1, c1, n_img, hw, groups, eps, gamma, beta, silu, stats, y):
        self._groupnorm(*map(_val, (x0, c0, x1, c1, n_img, hw, groups, eps, gamma, beta, silu, y)))

    def groupnorm_fused(self, x0, c0, part0, x1, c1, part1, n_img, hw, groups, eps, gamma, beta, silu, stats, y):
        # statistics recomputed from the data: the partial tables are checked by checking their producers' outputs
        self._groupnorm(*map(_val, (x0, c0, x1, c1, n_img, hw, groups, eps, gamma, beta, silu, y)))

    def layernorm(self, x, y, rows, c, eps, gamma, beta):
        x, y, rows, c, eps = map(_val, (x, y, rows, c, eps))
        out = F.layer_norm(_mem(x, rows * c, BF16).view(rows, c).float(), (c,), _mem(_val(gamma), c, F32),
                           _mem(_val(beta), c, F32), eps)
        _mem(y, rows * c, BF16).view(rows, c).copy_(out.to(BF16))

    def softmax_rows(self, x, rows, cols, ld, scale):
        v = _rows(_val(x), _val(rows), _val(cols), _val(ld))
        v.copy_(torch.softmax(v.float() * _val(scale), dim=-1).to(BF16))

    # ------------------------------------------------------------------------------------------------ layout / time path
    def nchw_to_nhwc8(self, x, dtype, n_img, c, hw, dup, y):
        x, dtype, n_img, c, hw, dup, y = map(_val, (x, dtype, n_img, c, hw, dup, y))
        src = _mem(x, n_img * c * hw, F32 if dtype == 0 else BF16).view(n_img, c, hw).float()
        n = n_img * (2 if dup else 1)
        out = torch.zeros(n, hw, 8)
        out[:n_img, :, :c] = src.permute(0, 2, 1)
        if dup:                                                 # classifier-free guidance: [latents | latents]
            out[n_img:] = out[:n_img]
        _mem(y, n * hw * 8, BF16).view(n, hw, 8).copy_(out.to(BF16))

    def nhwc_to_nchw(self, x, ld, n_img, c, hw, y, dtype):
        x, ld, n_img, c, hw, y, dtype = map(_val, (x, ld, n_img, c, hw, y, dtype))
        src = _rows(x, n_img * hw, c, ld).float().view(n_img, hw, c).permute(0, 2, 1)
        out_dtype = F32 if dtype == 0 else BF16
        _mem(y, n_img * c * hw, out_dtype).view(n_img, c, hw).copy_(src.to(out_dtype))

    def im2col3x3(self, x, y, n_img, h, w, c, stride):
        x, y, n_img, h, w, c, stride = map(_val, (x, y, n_img, h, w, c, stride))
        assert stride == 1
        src = F.pad(_mem(x, n_img * h * w * c, BF16).view(n_img, h, w, c), (0, 0, 1, 1, 1, 1))
        cols = [src[:, ky:ky + h, kx:kx + w, :] for ky in range(3) for kx in range(3)]     # tap-major, then channel
        _mem(y, n_img * h * w * 9 * c, BF16).view(n_img, h, w, 9 * c).copy_(torch.cat(cols, dim=-1))

    def timestep_embedding(self, t_dev, dim, out):
        t, dim = float(_mem(_val(t_dev), 1, F32)[0]), _val(dim)
        half = dim // 2
        f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=F32) / half)
        _mem(_val(out), dim, F32).copy_(torch.cat([torch.cos(t * f), torch.sin(t * f)]))

    def gemv(self, n_jobs, w, bias, add, y, n_arr, x, k, silu_in):
        n_jobs, k, silu_in = _val(n_jobs), _val(k), _val(silu_in)
        xv = _mem(_val(x), k, F32).clone()
        if silu_in:
            xv = F.silu(xv)
        for j in range(n_jobs):
            n = n_arr[j]
            out = _mem(w[j], n * k, BF16).view(n, k).float() @ xv
            if bias is not None and bias[j]:
                out = out + _mem(bias[j], n, F32)
            if add is not None and add[j]:
                out = out + _mem(add[j], n, F32)
            _mem(y[j], n, F32).copy_(out)
