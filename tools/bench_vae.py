"""VAE decode of 16 latents (512x512 images): native engine vs the PyTorch module (library kernels, bf16)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle.vae import make_vae
from sonicdiffusionbayeslab_b200.vae_engine import VaeEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
mod = make_vae(29, dtype=torch.bfloat16, device=dev)
eng = VaeEngine({k: v.detach() for k, v in mod.state_dict().items()}, n_img=B, device=dev)
z = torch.randn(B, 4, 64, 64, device=dev).bfloat16()


def timeit(f, reps=3):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        f()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


t_native = timeit(lambda: eng.decode(z))
with torch.no_grad():
    t_torch = timeit(lambda: mod.decode(z))
n_launch, flops = eng.stats()
print(f"VAE decode batch {B}: native {t_native:.1f} ms ({flops / t_native / 1e9:.0f} TFLOP/s, {n_launch} launches, "
      f"arena {eng.arena.bytes / 2**30:.2f} GiB) | torch bf16 {t_torch:.1f} ms | speed-up {t_torch / t_native:.2f}x")
if len(sys.argv) > 2:
    import ctypes
    from collections import defaultdict
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        prof = eng.plan.profile(ctypes.c_void_p(s.cuda_stream))
    torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for (kind, ms, fl), desc in zip(prof, eng.plan.log):
        a = agg[desc]; a[0] += 1; a[1] += ms; a[2] += fl
    for desc, (cnt, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        tf = f"{fl / ms / 1e9:7.1f} TF/s" if fl else "            "
        print(f"{ms:8.3f} ms x{cnt:<3d} {tf}  {desc}")
