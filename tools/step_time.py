"""A/B metric for kernel experiments: CUDA-graph replay time of one full UNet step at the bench batch.
Short bursts (3 replays, idle gaps: burst clocks) and one sustained run (40 replays: the power-capped clocks of the
timed loop).  Usage: [SONIC_LIB=...] python tools/step_time.py [tag] [n_latents]"""
import os
import statistics
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200.unet_engine import UNetEngine
from sonicdiffusionbayeslab_b200.unet_spec import random_unet_state_dict

tag = sys.argv[1] if len(sys.argv) > 1 else ""
n_lat = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
torch.manual_seed(0)                      # same inputs in every process: the digest below is comparable across A/B runs
eng = UNetEngine(random_unet_state_dict(29), n_latents=n_lat, cfg_dup=True, device=dev)
eng.x_in.normal_()
eng.set_context(torch.randn(2 * n_lat, 77, 768, device=dev).bfloat16())
eng.capture_graphs()


def run(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        eng.forward(500.0)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


run(3)
bursts = []
for _ in range(9):
    time.sleep(0.3)
    bursts.append(run(3))
sustained = run(40)
digest = zlib.crc32(eng.forward(500.0).float().cpu().numpy().tobytes())
print(f"{tag} UNet step (graph replay, UNet batch {2 * n_lat}): burst min {min(bursts):.3f} median "
      f"{statistics.median(bursts):.3f} ms; sustained x40 {sustained:.3f} ms; eps crc32 {digest:08x}")
