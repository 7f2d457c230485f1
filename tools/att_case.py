"""One attention case against torch SDPA: python tools/att_case.py B H Sq Sk d [causal]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from sonicdiffusionbayeslab_b200 import kernels as k
B, H, Sq, Sk, d = map(int, sys.argv[1:6])
causal = len(sys.argv) > 6 and sys.argv[6] == "1"
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
C = H * d
q = torch.randn(B * Sq, C, device=dev, generator=g).bfloat16()
kk = torch.randn(B * Sk, C, device=dev, generator=g).bfloat16()
v = torch.randn(B * Sk, C, device=dev, generator=g).bfloat16()
out = k.attention(q, kk, v, batch=B, heads=H, seq_q=Sq, seq_k=Sk, head_dim=d, causal=causal)
torch.cuda.synchronize()
qf, kf, vf = (x.float().reshape(B, -1, H, d).transpose(1, 2) for x in (q, kk, v))
ref = F.scaled_dot_product_attention(qf, kf, vf, is_causal=causal).transpose(1, 2).reshape(B * Sq, C)
print(f"var {os.environ.get('SONIC_ATT_VAR')} attn B{B} H{H} Sq{Sq} Sk{Sk} d{d}: max_abs={(out.float() - ref).abs().max().item():.3e}", flush=True)
