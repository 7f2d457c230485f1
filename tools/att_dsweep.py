import os, sys
sys.path.insert(0, "/root/repo")
import torch
from sonicdiffusionbayeslab_b200 import kernels as k
dev = torch.device("cuda:0")
def bench(B, H, Sq, Sk, d, reps=5):
    C = H * d
    q = torch.randn(B * Sq, C, device=dev).bfloat16()
    kv = torch.randn(B * Sk, 2 * C, device=dev).bfloat16()
    f = lambda: k.attention(q, kv[:, :C], kv[:, C:], batch=B, heads=H, seq_q=Sq, seq_k=Sk, head_dim=d)
    for _ in range(2): f()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): f()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print(f"var {os.environ.get('SONIC_ATT_VAR')} d{d}: {ms:.3f} ms", flush=True)
for var in ("0",):
    os.environ["SONIC_ATT_VAR"] = var
    for d in (16, 32, 40, 48, 64):
        bench(32, 8, 4096, 4096, d)
