"""Thin torch-tensor wrappers over the libsonic C ABI (one function per entry point).

These exist for the unit/parity tests and for host code that needs a single operator; the
sampling engine itself calls ``sonic_unet_*`` / ``sonic_latent_update`` which run whole
launch plans natively.  Tensors are only used as device-memory handles (``data_ptr``).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import GemmArgs, check, lib, ptr, stream_ptr

EPI_NONE, EPI_GEGLU = 0, 1


def _bf16c(t: torch.Tensor) -> torch.Tensor:
    assert t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous(), (t.dtype, t.device, t.is_contiguous())
    return t


def pack_conv3x3_weight(w_oihw: torch.Tensor) -> torch.Tensor:
    """OIHW -> [tap=kh*3+kw][O][I] bf16 (K-major per tap), the layout the TMA B-box reads."""
    O, I, kh, kw = w_oihw.shape
    return w_oihw.permute(2, 3, 0, 1).reshape(kh * kw, O, I).contiguous().to(torch.bfloat16)


def pack_geglu(w: torch.Tensor, b: torch.Tensor, block_n: int):
    """Interleave the value / gate halves of ff.net.0.proj per ``block_n`` tile."""
    n2, K = w.shape
    n = n2 // 2
    half = block_n // 2
    assert n % half == 0
    wv, wg = w[:n].reshape(n // half, half, K), w[n:].reshape(n // half, half, K)
    wp = torch.cat([wv, wg], dim=1).reshape(n2, K).contiguous()
    bp = torch.cat([b[:n].reshape(-1, half), b[n:].reshape(-1, half)], dim=1).reshape(n2).contiguous()
    return wp, bp


def gemm_block_n(N: int, n_img: int, H: int, W: int, epilogue: int = EPI_NONE) -> int:
    return lib().sonic_gemm_block_n(N, n_img, H, W, epilogue)


def conv_gemm(a0, w, N, *, taps=1, n_img=1, H=1, W=None, c0=None, a1=None, c1=0, bias=None,
              row_bias=None, residual=None, out=None, epilogue=EPI_NONE, block_n=0):
    """out[M, N'] = epilogue(implicit_gemm(A, w)); see ``sonic_conv_gemm`` in include/sonic.h.

    ``a0`` / ``a1`` are NHWC bf16 tensors whose last dim is the pixel pitch; a Linear over
    [M, K] uses the defaults (n_img=1, H=1, W=M).
    """
    _bf16c(a0), _bf16c(w)
    ld0 = a0.shape[-1]
    c0 = ld0 if c0 is None else c0
    if W is None:
        W = a0.numel() // ld0
    M = n_img * H * W
    n_out = N // 2 if epilogue == EPI_GEGLU else N
    if out is None:
        out = torch.empty((M, n_out), device=a0.device, dtype=torch.bfloat16)
    args = GemmArgs()
    args.a0, args.c0, args.ld0 = a0.data_ptr(), c0, ld0
    if a1 is not None:
        _bf16c(a1)
        args.a1, args.c1, args.ld1 = a1.data_ptr(), c1 or a1.shape[-1], a1.shape[-1]
    args.n_img, args.H, args.W = n_img, H, W
    args.w, args.N, args.taps = w.data_ptr(), N, taps
    for name, t in (("bias", bias), ("row_bias", row_bias)):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous()
            setattr(args, name, t.data_ptr())
    if residual is not None:
        _bf16c(residual)
        args.residual, args.ld_res = residual.data_ptr(), residual.shape[-1]
    args.out, args.ld_out = out.data_ptr(), out.shape[-1]
    args.epilogue, args.block_n = epilogue, block_n
    check(lib().sonic_conv_gemm(C.byref(args), stream_ptr()), "sonic_conv_gemm")
    return out
