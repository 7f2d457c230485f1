"""Helper of test_gemm_pair_gpu.py: runs a fixed list of GEMM / GroupNorm cases on cuda:0 and prints one line per case,
`name sha256(output bytes)`.  The kernel variant is chosen by the environment of THIS process (SONIC_GEMM_PAIR,
SONIC_GN_CLUSTER are read once per process), so the test runs it twice and compares the lines."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200 import kernels as k

dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(11)


def bf(*shape, scale=1.0):
    return (scale * torch.randn(*shape, device=dev, generator=g)).bfloat16()


def digest(name, t):
    torch.cuda.synchronize()
    print(name, hashlib.sha256(t.contiguous().view(torch.uint8).cpu().numpy().tobytes()).hexdigest(), flush=True)


# plain / residual / odd tile counts (3 M tiles: the last pair has one out-of-range tile)
for M, N, K in ((4096, 320, 320), (384, 640, 704), (1000, 1280, 1280), (130, 64, 64)):
    a, w, res = bf(M, K), bf(N, K, scale=K ** -0.5), bf(M, N)
    digest(f"linear_{M}_{N}_{K}", k.conv_gemm(a, w, N, bias=torch.randn(N, device=dev, generator=g), residual=res))
# 3x3 convolution with per-image bias, concat source, GroupNorm partials
for B, H, W, C, Co in ((2, 64, 64, 320, 320), (3, 16, 16, 1280, 640), (3, 8, 8, 1280, 1280)):
    x = bf(B, H, W, C)
    w = bf(Co, C, 3, 3, scale=(9 * C) ** -0.5)
    part = k.gn_partial_buffer(B * H * W, Co, dev)
    out = k.conv_gemm(x, k.pack_conv3x3_weight(w), Co, taps=9, n_img=B, H=H, W=W,
                      row_bias=torch.randn(B, Co, device=dev, generator=g), gn_partial=part)
    digest(f"conv3x3_{B}_{H}_{C}_{Co}", out)
    digest(f"conv3x3_{B}_{H}_{C}_{Co}_gnpart", part)
# GEGLU (value half from the leader's B rows, gate half from the peer's)
for M, C in ((4096, 320), (640, 1280)):
    a = bf(M, C)
    w, b = bf(8 * C, C, scale=C ** -0.5), torch.randn(8 * C, device=dev, generator=g)
    bn = k.gemm_block_n(8 * C, 1, 1, M, k.EPI_GEGLU)
    wp, bp = k.pack_geglu(w, b, bn)
    digest(f"geglu_{M}_{C}", k.conv_gemm(a, wp, 8 * C, bias=bp, epilogue=k.EPI_GEGLU, block_n=bn))
# stride 2 and the phase form of the upsample convolution
x = bf(2, 32, 32, 320)
w = bf(320, 320, 3, 3, scale=(9 * 320) ** -0.5)
digest("conv_s2", k.conv_gemm(x, k.pack_conv3x3_weight(w), 320, taps=9, n_img=2, H=16, W=16, stride=2))
x = bf(2, 16, 16, 640)
w = bf(640, 640, 3, 3, scale=(9 * 640) ** -0.5)
digest("conv_up", k.conv_gemm(x, k.pack_upsample_conv_weight(w), 640, taps=9, n_img=2, H=16, W=16, upsample=1))
