"""Standalone launches of the shapes that bound the 64x64 transformer blocks, for `ncu --set full` (one profiled
launch per shape: everything before cudaProfilerStart is warm-up).

    ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_shapes \
        python tools/ncu_shapes.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200 import kernels as K

dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
M, C = 131072, 320


def bf(*shape, scale=1.0):
    return (scale * torch.randn(*shape, device=dev, generator=g)).bfloat16()


a = bf(M, C)
res = bf(M, C)
w = bf(C, C, scale=C ** -0.5)
bias = torch.zeros(C, device=dev)
bn = K.gemm_block_n(C, 1, 1, M)
stats, parts = K.ln_stats_buffer(M, C, bn, dev)
h = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
side = torch.zeros(M, 64, device=dev, dtype=torch.bfloat16)
rstd = torch.ones(M, device=dev)
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
w_qkv = K.fold_layernorm(torch.randn(3 * C, C, device=dev, generator=g) * C ** -0.5, None, gamma, beta)
wg = K.fold_layernorm(torch.randn(8 * C, C, device=dev, generator=g) * C ** -0.5, torch.zeros(8 * C, device=dev), gamma, beta)
bng = K.gemm_block_n(8 * C, 1, 1, M, K.EPI_GEGLU)
w_geglu, _ = K.pack_geglu(wg, torch.zeros(8 * C, device=dev), bng)
qkv = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
ff = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
x4 = bf(32, 64, 64, C)
w3 = K.pack_conv3x3_weight(bf(C, C, 3, 3, scale=(9 * C) ** -0.5))
kv = bf(32 * 77, 2 * C)
q = bf(M, C)
ao = torch.empty(M, C, device=dev, dtype=torch.bfloat16)

jobs = [
    ("proj +res +lnstats K=320", lambda: K.conv_gemm(a, w, C, bias=bias, residual=res, out=h, block_n=bn, ln_stats_out=stats)),
    ("ln_side", lambda: K.ln_side(stats, C, side=side, rstd=rstd)),
    ("qkv lnfold N=960", lambda: K.conv_gemm(h, w_qkv, 3 * C, a1=side, out=qkv, row_scale=rstd)),
    ("geglu lnfold N=2560", lambda: K.conv_gemm(h, w_geglu, 8 * C, a1=side, out=ff, epilogue=K.EPI_GEGLU, block_n=bng, row_scale=rstd)),
    ("conv3x3 320", lambda: K.conv_gemm(x4, w3, C, taps=9, n_img=32, H=64, W=64, bias=bias, out=h)),
    ("cross-attention 4096x77 d40", lambda: K.attention(q, kv[:, :C], kv[:, C:], batch=32, heads=8, seq_q=4096, seq_k=77, head_dim=40, out=ao)),
]
for name, fn in jobs:
    for _ in range(3):
        fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for name, fn in jobs:
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", [n for n, _ in jobs])
