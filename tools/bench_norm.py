"""Times the HBM-bound kernels at the bench shapes (UNet batch 32) against the measured copy bandwidth."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200 import kernels as k

dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6540.5
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(f, reps=5):
    f()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        f()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts)


for rows, C in [(131072, 320), (32768, 640), (8192, 1280)]:
    x = torch.randn(rows, C, device=dev).bfloat16()
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    y = torch.empty_like(x)
    ms = timeit(lambda: k.layernorm(x, g, b, out=y))
    ref = torch.nn.functional.layer_norm(x.float(), (C,), g, b)
    err = (y.float() - ref).abs().max().item()
    gb = 2 * x.numel() * 2 / 1e9
    print(f"layernorm rows={rows} C={C}: {ms * 1e3:.1f} us  {gb / ms * 1e3:.0f} GB/s  ({gb / ms * 1e3 / peak:.0%} of measured copy) err={err:.3e}")
    n_img = 32
    hw = rows // n_img
    for silu in (True, False):
        ms = timeit(lambda: k.groupnorm(x, g, b, n_img=n_img, hw=hw, silu=silu, out=y))
        gb3 = 3 * x.numel() * 2 / 1e9
        print(f"groupnorm rows={rows} C={C} silu={int(silu)}: {ms * 1e3:.1f} us  {gb3 / ms * 1e3:.0f} GB/s algorithmic (2 reads + 1 write; "
              f"{gb3 / ms * 1e3 / peak:.0%} of measured copy)")
