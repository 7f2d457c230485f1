"""Per-operator CUDA-event profile of the UNet plan at the bench batch (eager run, not a graph)."""
import ctypes
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200.unet_engine import UNetEngine
from sonicdiffusionbayeslab_b200.unet_spec import random_unet_state_dict

n_lat = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
eng = UNetEngine(random_unet_state_dict(29), n_latents=n_lat, cfg_dup=True, device=dev)
eng.x_in.normal_()
eng.set_context(torch.randn(2 * n_lat, 77, 768, device=dev).bfloat16())
plan = eng.plans[sys.argv[2] if len(sys.argv) > 2 else "full"]
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(2):
        prof = plan.profile(ctypes.c_void_p(s.cuda_stream))
torch.cuda.synchronize()
assert len(prof) == len(plan.log), (len(prof), len(plan.log))
agg = defaultdict(lambda: [0, 0.0, 0.0])
tot = sum(ms for _, ms, _ in prof)
for (kind, ms, fl), desc in zip(prof, plan.log):
    a = agg[desc]
    a[0] += 1; a[1] += ms; a[2] += fl
print(f"total {tot:.2f} ms over {len(prof)} ops, arena {eng.arena.bytes / 2**30:.2f} GiB")
for desc, (cnt, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tf = f"{fl / ms / 1e9:7.1f} TF/s" if fl else "            "
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  x{cnt:<3d} {tf}  {desc}")
