"""Timeline of CTA 0 of one conv_gemm launch (needs libsonic built with -DSONIC_GEMM_TRACE).
Usage: python tools/gemm_trace.py M N K [res] [geglu]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200 import kernels as k
from sonicdiffusionbayeslab_b200._lib import lib

M, N, K = (int(a) for a in sys.argv[1:4])
use_res = "res" in sys.argv
geglu = "geglu" in sys.argv
dev = torch.device("cuda:0")
a = torch.randn(M, K, device=dev).bfloat16()
w = torch.randn(N, K, device=dev).bfloat16() * 0.05
bias = torch.randn(N, device=dev)
epi, bn = k.EPI_NONE, 0
if geglu:
    bn = k.gemm_block_n(N, 1, 1, M, k.EPI_GEGLU)
    w, bias = k.pack_geglu(w, bias, bn)
    epi = k.EPI_GEGLU
res = torch.randn(M, N, device=dev).bfloat16() if use_res else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()
    out = k.conv_gemm(a, w, N, bias=bias, residual=res, epilogue=epi, block_n=bn)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush.zero_()
s.record()
out = k.conv_gemm(a, w, N, bias=bias, residual=res, epilogue=epi, block_n=bn)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e)
print(f"M={M} N={N} K={K} res={use_res}: {ms * 1e3:.1f} us  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s")
buf = (C.c_longlong * (4 * 64 * 4))()
assert lib().sonic_debug_gemm_trace(buf) == 0
tr = torch.tensor(list(buf)).view(4, 64, 4)
t0 = int(tr[0, 0, 0])
print("tile | TMA: start  issued | MMA: start acc_free first_full committed | EPI: start res_issued acc_full done | chunk0: ld_issued math_done smem_written store_issued")
for t in range(0, 14):
    r = [f"{t:3d} |"] + [f"{int(tr[0, t, i]) - t0:8d}" for i in range(2)] + ["|"] + \
        [f"{int(tr[1, t, i]) - t0:8d}" for i in range(4)] + ["|"] + [f"{int(tr[2, t, i]) - t0:8d}" for i in range(4)] + ["|"] + [f"{int(tr[3, t, i]) - t0:8d}" for i in range(4)]
    print(" ".join(r))
