"""Bit-level A/B of a launch-side or kernel-side switch across EVERY native engine: CRC32 digests of the outputs of the
UNet (UNet batch 32 eager + CUDA-graph replays, DeepCache cached plan, UNet batch 2), the VAE decoder and both CLIP
towers on seeded inputs, plus the UNet step time.  All kernels are deterministic, so two processes that differ only in
a switch that must not change results (SONIC_PDL, SONIC_GEMM_PAIR, suspend hints ...) print identical digest lines;
within one process every replay must reproduce the first digest (a race shows up as a second distinct value).

    python tools/engine_digest.py base;  SONIC_PDL=1 python tools/engine_digest.py pdl
"""
import os
import statistics
import sys
import time
import warnings
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200.unet_engine import PackedWeights, UNetEngine
from sonicdiffusionbayeslab_b200.unet_spec import random_unet_state_dict

tag = sys.argv[1] if len(sys.argv) > 1 else ""
replays = int(sys.argv[2]) if len(sys.argv) > 2 else 24
dev = torch.device("cuda:0")
torch.manual_seed(0)


def crc(t):
    return f"{zlib.crc32(t.float().cpu().numpy().tobytes()):08x}"


def line(what, digests, extra=""):
    distinct = sorted(set(digests))
    print(f"[{tag}] {what}: {distinct[0]} x{len(digests)}" + (f"  !! {len(distinct)} DISTINCT: {distinct}" if
                                                                len(distinct) > 1 else "") + extra, flush=True)
    return len(distinct) == 1


ok = True
w = PackedWeights(random_unet_state_dict(29), dev)
g = torch.Generator(device="cuda").manual_seed(29)
lat = torch.randn(16, 4, 64, 64, device=dev, generator=g).bfloat16()
ctx = torch.randn(32, 77, 768, device=dev, generator=g).bfloat16()

# ---- UNet, UNet batch 32 (the bench shape): eager, graph replays, DeepCache cached plan
eng = UNetEngine(w, n_latents=16, cfg_dup=True, device=dev, cache_branch=0)
eng.x_in.copy_(lat)
eng.set_context(ctx)
ok &= line("unet32 eager", [crc(eng.forward(481.0)) for _ in range(3)])
eng.capture_graphs()
ok &= line("unet32 graph", [crc(eng.forward(481.0)) for _ in range(replays)])
if "cached" in eng.plans:
    ok &= line("unet32 deepcache cached plan", [crc(eng.forward(461.0, cached=True)) for _ in range(6)])


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
for _ in range(7):
    time.sleep(0.3)
    bursts.append(run(3))
sustained = run(40)
print(f"[{tag}] unet32 step (graph replay): burst min {min(bursts):.3f} median {statistics.median(bursts):.3f} ms; "
      f"sustained x40 {sustained:.3f} ms", flush=True)
ok &= line("unet32 graph after the timing loop", [crc(eng.forward(481.0)) for _ in range(4)])

# ---- UNet batch 2 (other tile schedules, one-tile attention, unpaired GEMMs)
small = UNetEngine(w, n_latents=1, cfg_dup=True, device=dev)
small.x_in.copy_(lat[3:4])
small.set_context(torch.stack([ctx[3], ctx[19]]))
ok &= line("unet2 eager", [crc(small.forward(481.0)) for _ in range(6)])
small.capture_graphs()
ok &= line("unet2 graph", [crc(small.forward(481.0)) for _ in range(6)])
del eng, small

# ---- VAE decoder (two-kernel GroupNorm form, 512-channel 1-head attention, upsample convolutions)
from sonicdiffusionbayeslab_b200.vae_engine import VaeEngine
from sonicdiffusionbayeslab_b200.vae_spec import random_vae_state_dict

vae = VaeEngine(random_vae_state_dict(29), n_img=2, latent=64, device=dev)
z = torch.randn(2, 4, 64, 64, device=dev, generator=g).bfloat16()
ok &= line("vae decode 2x512x512", [crc(vae.decode(z)) for _ in range(4)])
del vae

# ---- CLIP towers (LayerNorm kernels, QuickGELU epilogue, causal attention, 1-D GEMMs)
from sonicdiffusionbayeslab_b200.clip_engine import ClipTextEngine, ClipVisionEngine
from sonicdiffusionbayeslab_b200.metrics.metrics import make_clip_model

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    weights, tok = make_clip_model(None)
sd = {k: v.detach() for k, v in weights.state_dict().items()}
n = 4
vis, txt = ClipVisionEngine(sd, n=n, device=dev), ClipTextEngine(sd, n=n, device=dev)
img = (torch.rand(n, 3, 512, 512, device=dev, generator=g) * 255).to(torch.uint8)
ids, _ = tok(["a photo of a cat", "two dogs running on the beach at sunset", "x", "a " * 60])
ok &= line("clip image features (uint8 512x512 in)", [crc(vis.image_features_from_images(img)) for _ in range(4)])
ok &= line("clip text features", [crc(txt.text_features(ids.to(dev))) for _ in range(4)])
print(f"[{tag}] {'ALL DETERMINISTIC' if ok else 'NON-DETERMINISTIC OUTPUT'}", flush=True)
sys.exit(0 if ok else 1)
