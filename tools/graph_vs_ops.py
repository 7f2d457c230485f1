"""How much of a UNet step is NOT inside kernels: CUDA-graph replay time of the plan (short bursts, so the clocks are
the same as for the per-operator profile) against the sum of the per-operator CUDA-event times of an eager run."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200.unet_engine import UNetEngine
from sonicdiffusionbayeslab_b200.unet_spec import random_unet_state_dict

n_lat = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
eng = UNetEngine(random_unet_state_dict(29), n_latents=n_lat, cfg_dup=True, device=dev)
eng.x_in.normal_()
eng.set_context(torch.randn(2 * n_lat, 77, 768, device=dev).bfloat16())
plan = eng.plans["full"]
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(2):
        prof = plan.profile(ctypes.c_void_p(s.cuda_stream))
torch.cuda.synchronize()
ops = sum(ms for _, ms, _ in prof)
eng.capture_graphs()
for reps in (1, 3):
    for _ in range(2):
        eng.forward(500.0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        eng.forward(500.0)
    b.record()
    torch.cuda.synchronize()
    print(f"graph replay x{reps}: {a.elapsed_time(b) / reps:.3f} ms per step; per-operator events sum {ops:.3f} ms over {len(prof)} ops")
