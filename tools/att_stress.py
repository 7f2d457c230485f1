"""Stress the attention kernels for rare races: many repetitions of ragged / multi-wave shapes against torch SDPA."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from sonicdiffusionbayeslab_b200 import kernels as k

dev = torch.device("cuda:0")
cases = [(4, 8, 4096, 4096, 40, False), (32, 8, 4096, 77, 40, False), (8, 8, 1024, 1024, 80, False), (16, 12, 197, 197, 64, False),
         (3, 5, 1152, 100, 64, False), (4, 8, 512, 512, 40, True), (32, 8, 256, 77, 160, False), (2, 3, 2048, 128, 48, False),
         (6, 4, 640, 640, 40, False), (32, 12, 77, 77, 64, True)]
worst = {}
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 25
for r in range(reps):
    for (B, H, Sq, Sk, d, causal) in cases:
        g = torch.Generator(device="cuda").manual_seed(1000 * r + Sq + Sk + d)
        C = H * d
        q = torch.randn(B * Sq, C, device=dev, generator=g).bfloat16()
        kk = torch.randn(B * Sk, C, device=dev, generator=g).bfloat16()
        v = torch.randn(B * Sk, C, device=dev, generator=g).bfloat16()
        outs = [k.attention(q, kk, v, batch=B, heads=H, seq_q=Sq, seq_k=Sk, head_dim=d, causal=causal) for _ in range(3)]
        torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), ("non-deterministic", B, H, Sq, Sk, d, causal, r)
        if r % 5 == 0:
            qf, kf, vf = (x.float().reshape(B, -1, H, d).transpose(1, 2) for x in (q, kk, v))
            ref = F.scaled_dot_product_attention(qf, kf, vf, is_causal=causal).transpose(1, 2).reshape(B * Sq, C)
            err = ((outs[0].float() - ref).abs().max() / ref.abs().max()).item()
            worst[(B, H, Sq, Sk, d, causal)] = max(worst.get((B, H, Sq, Sk, d, causal), 0.0), err)
            assert err < 2e-2, (err, B, H, Sq, Sk, d, causal, r)
print("ok", reps, "repetitions; worst relative errors:")
for c, e in worst.items():
    print(" ", c, f"{e:.3e}")
