"""CUDA-event timing of the UNet's attention shapes at UNet batch 32, each alone (10 launches after 3 warm-ups)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200 import kernels as K

dev = torch.device("cuda:0")
tag = sys.argv[1] if len(sys.argv) > 1 else ""
B, H = 32, 8
tot = 0.0
for Sq, Sk, d in ((4096, 4096, 40), (4096, 77, 40), (1024, 1024, 80), (1024, 77, 80), (256, 256, 160), (256, 77, 160),
                  (64, 64, 160), (64, 77, 160)):
    C = H * d
    q = torch.randn(B * Sq, C, device=dev).bfloat16()
    kv = torch.randn(B * Sk, 2 * C, device=dev).bfloat16()
    out = torch.empty(B * Sq, C, device=dev, dtype=torch.bfloat16)
    fn = lambda: K.attention(q, kv[:, :C], kv[:, C:], batch=B, heads=H, seq_q=Sq, seq_k=Sk, head_dim=d, out=out)   # noqa: E731
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / 10 * 1e3
    tot += us if Sk == 77 else 0
    print(f"{tag} {us:8.1f} us  {4.0 * B * H * Sq * Sk * d / us / 1e6:7.1f} TF/s  Sq={Sq} Sk={Sk} d={d}")
print(f"{tag} cross-attention total {tot:.1f} us")
