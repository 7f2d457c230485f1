"""ncu target: the 64x64 self-attention layer at the bench batch (B=32, H=8, S=4096, d=40), last call profiled."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200 import kernels as k

B, H, S, d = 32, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device("cuda:0")
C = H * d
qkv = torch.randn(B * S, 3 * C, device=dev).bfloat16()
f = lambda: k.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], batch=B, heads=H, seq_q=S, seq_k=S, head_dim=d)
for _ in range(2):
    f()
torch.cuda.synchronize()
torch.cuda.profiler.start()
f()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
