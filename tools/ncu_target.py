"""Short ncu target: eager UNet forwards at the bench batch (UNet batch 32), no CUDA graph.  Only the
LAST forward sits inside the cudaProfilerStart/Stop range, so with ``--profile-from-start off`` ncu sees
exactly the launches of one warm UNet step (none of torch's weight-initialisation kernels).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/ncu_target.py
    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:conv_gemm_kernel -c 12 -o gpurun_out/prof_gemm python tools/ncu_target.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200.unet_engine import UNetEngine
from sonicdiffusionbayeslab_b200.unet_spec import random_unet_state_dict

n_lat = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
eng = UNetEngine(random_unet_state_dict(29), n_latents=n_lat, cfg_dup=True, device=dev, build_cached=False)
eng.x_in.normal_()
eng.set_context(torch.randn(2 * n_lat, 77, 768, device=dev).bfloat16())
for _ in range(2):
    eng.forward(500.0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.forward(500.0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", eng.stats("full"))
