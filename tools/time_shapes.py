"""CUDA-event timing of the GEMM shapes of the 64x64 / 32x32 transformer blocks, each alone, back to back
(20 launches after 5 warm-ups).  For A/B experiments on the epilogue:  python tools/time_shapes.py [tag]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200 import kernels as K

dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)


def bf(*shape, scale=1.0):
    return (scale * torch.randn(*shape, device=dev, generator=g)).bfloat16()


def timed(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


rows = []
for M, C in ((131072, 320), (32768, 640), (8192, 1280)):
    a, res = bf(M, C), bf(M, C)
    w = bf(C, C, scale=C ** -0.5)
    bias = torch.zeros(C, device=dev)
    bn = K.gemm_block_n(C, 1, 1, M)
    stats, parts = K.ln_stats_buffer(M, C, bn, dev)
    h = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    side, rstd = torch.zeros(M, 64, device=dev, dtype=torch.bfloat16), torch.ones(M, device=dev)
    gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    w_q = K.fold_layernorm(torch.randn(C, C, device=dev, generator=g) * C ** -0.5, None, gam, bet)
    w_qkv = K.fold_layernorm(torch.randn(3 * C, C, device=dev, generator=g) * C ** -0.5, None, gam, bet)
    wg = K.fold_layernorm(torch.randn(8 * C, C, device=dev, generator=g) * C ** -0.5, torch.zeros(8 * C, device=dev), gam, bet)
    bng = K.gemm_block_n(8 * C, 1, 1, M, K.EPI_GEGLU)
    w_geglu, _ = K.pack_geglu(wg, torch.zeros(8 * C, device=dev), bng)
    w_ff2 = bf(C, 4 * C, scale=(4 * C) ** -0.5)
    qkv = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
    ff = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
    q = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    part = K.gn_partial_buffer(M, C, dev)
    jobs = [
        (f"proj  M={M} N={C} K={C} +res +lnstats (bn {bn})", 2.0 * M * C * C,
         lambda: K.conv_gemm(a, w, C, bias=bias, residual=res, out=h, block_n=bn, ln_stats_out=stats)),
        (f"proj  M={M} N={C} K={C} +res +gnstats", 2.0 * M * C * C,
         lambda: K.conv_gemm(a, w, C, bias=bias, residual=res, out=h, gn_partial=part)),
        (f"proj  M={M} N={C} K={C} plain", 2.0 * M * C * C, lambda: K.conv_gemm(a, w, C, bias=bias, out=h)),
        (f"to_q  M={M} N={C} K={C} lnfold", 2.0 * M * C * C, lambda: K.conv_gemm(a, w_q, C, a1=side, out=q, row_scale=rstd)),
        (f"qkv   M={M} N={3 * C} K={C} lnfold", 6.0 * M * C * C, lambda: K.conv_gemm(a, w_qkv, 3 * C, a1=side, out=qkv, row_scale=rstd)),
        (f"geglu M={M} N={8 * C} K={C} lnfold", 16.0 * M * C * C,
         lambda: K.conv_gemm(a, w_geglu, 8 * C, a1=side, out=ff, epilogue=K.EPI_GEGLU, block_n=bng, row_scale=rstd)),
        (f"ff2   M={M} N={C} K={4 * C} +res", 8.0 * M * C * C, lambda: K.conv_gemm(ff, w_ff2, C, bias=bias, residual=res, out=h)),
    ]
    for name, fl, fn in jobs:
        us = timed(fn)
        rows.append((name, us, fl / us / 1e6))
    del a, res, qkv, ff
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for name, us, tf in rows:
    print(f"{tag} {us:8.1f} us {tf:7.1f} TF/s  {name}")
print(f"{tag} total {sum(r[1] for r in rows):.1f} us")
