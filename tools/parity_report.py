"""Measured ABSOLUTE per-step latent errors on the unit-variance fixture:
engine (bf16 and fp32 latent I/O) and stock-PyTorch bf16, each teacher-forced from the fp32 oracle.

    python tools/parity_report.py > profiles/r2_parity_abs.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import parity_lib as PL

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
net, net16, scale = PL.unit_variance_unet(dev)
sd = dict(net.state_dict())
print(f"unit-variance fixture: conv_out scaled by {scale:.4f}")
r = PL.teacher_forced("dpmpp25", net, None, sd, dev, io_dtype=torch.bfloat16, B=16, max_steps=3)
print(f"batch 16 (UNet batch 32) dpmpp25 steps 1-3 vs fp32 oracle: engine max-abs {r['engine']}  |x|max {max(r['xmax']):.2f}")
for name in PL.CASES:
    for io in (torch.bfloat16, torch.float32):
        r = PL.teacher_forced(name, net, net16, sd, dev, io_dtype=io)
        e, f = r["engine"], r["torch_bf16"]
        print(f"{name:22s} io={str(io).split('.')[-1]:8s} steps={len(e):2d} |x|max={max(r['xmax']):5.2f}  "
              f"engine max-abs: worst {max(e):.3e} median {sorted(e)[len(e) // 2]:.3e}"
              + (f"  | torch-bf16: worst {max(f):.3e} median {sorted(f)[len(f) // 2]:.3e}" if f else ""))
        print("    per step: " + " ".join(f"{v:.1e}" for v in e))
