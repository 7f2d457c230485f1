"""Head-dim sweep of the self-attention kernels at a fixed (batch, heads, tokens): separates the cost of the head dim
(MMA K steps of S, MMA N of PV, smem atoms) from the softmax.  Usage: python tools/att_dsweep.py [S] [d ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200 import kernels as k

dev = torch.device("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dims = [int(a) for a in sys.argv[2:]] or [40, 48, 64, 72, 80, 96, 128, 136, 160]


def bench(B, H, Sq, Sk, d, reps=10):
    C = H * d
    q = torch.randn(B * Sq, C, device=dev).bfloat16()
    kv = torch.randn(B * Sk, 2 * C, device=dev).bfloat16()
    f = lambda: k.attention(q, kv[:, :C], kv[:, C:], batch=B, heads=H, seq_q=Sq, seq_k=Sk, head_dim=d)   # noqa: E731
    for _ in range(3):
        f()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        f()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / reps * 1e3
    print(f"{os.environ.get('SONIC_ATT2_WIDE', '1')} S={Sq} d={d:3d}: {us:8.1f} us  {4.0 * B * H * Sq * Sk * d / us / 1e6:7.1f} TF/s", flush=True)


for d in dims:
    bench(32, 8, S, S, d)
