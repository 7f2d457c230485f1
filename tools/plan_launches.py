"""Kernel launches per UNet plan as the plan itself counts them (what bench.py reports as gpu_launches); compare with
the ncu launch list of the same build (profiles/*_launch_list_summary.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200.unet_engine import UNetEngine
from sonicdiffusionbayeslab_b200.unet_spec import random_unet_state_dict

dev = torch.device("cuda:0")
for n_lat in (16, 1):
    eng = UNetEngine(random_unet_state_dict(29), n_latents=n_lat, cfg_dup=True, device=dev)
    print(f"UNet batch {2 * n_lat}:", {name: eng.stats(name)[0] for name in eng.plans})
