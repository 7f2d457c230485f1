"""GPU debugging aid: engine UNet forward vs the oracle UNet (fp32 on the GPU), full size."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle.unet import make_unet
from sonicdiffusionbayeslab_b200.unet_engine import UNetEngine

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
n_lat = int(sys.argv[1]) if len(sys.argv) > 1 else 1
net = make_unet(29).to(dev)
sd = {k: v for k, v in net.state_dict().items()}
t0 = time.time()
eng = UNetEngine(sd, n_latents=n_lat, cfg_dup=True, io_dtype=torch.float32, device=dev)
torch.cuda.synchronize()
print(f"engine built in {time.time() - t0:.1f}s; arena {eng.arena.bytes / 2**20:.0f} MiB; plan stats full={eng.stats('full')} "
      f"cached={eng.stats('cached')}", flush=True)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(n_lat, 4, 64, 64, device=dev, generator=g)
ctx = torch.randn(2 * n_lat, 77, 768, device=dev, generator=g)
ctx_bf = ctx.bfloat16()
eng.x_in.copy_(x)
eng.set_context(ctx_bf)
for t in (951.0, 501.0, 1.0):
    eps = eng.forward(t).clone()
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = net(torch.cat([x, x]), torch.tensor(t, device=dev), ctx_bf.float())[0]
    err = (eps - ref).abs()
    print(f"t={t}: max_abs={err.max().item():.4e} mean_abs={err.mean().item():.4e} ref_absmax={ref.abs().max().item():.3f} "
          f"ref_std={ref.std().item():.3f} nan={torch.isnan(eps).any().item()}", flush=True)

# DeepCache cached step vs oracle DeepCache semantics
from oracle.deepcache import deepcache_forward  # noqa: E402

with torch.no_grad():
    state = {}
    ref_full = deepcache_forward(net, torch.cat([x, x]), torch.tensor(951.0, device=dev), ctx_bf.float(), state, full=True)
    x2 = x * 0.9 + 0.1
    ref_c = deepcache_forward(net, torch.cat([x2, x2]), torch.tensor(913.0, device=dev), ctx_bf.float(), state, full=False)
eng.x_in.copy_(x)
eng.forward(951.0)
eng.x_in.copy_(x2)
eps_c = eng.forward(913.0, cached=True).clone()
torch.cuda.synchronize()
print(f"deepcache cached step: max_abs={(eps_c - ref_c).abs().max().item():.4e} ref_absmax={ref_c.abs().max().item():.3f}")

# timing: eager plan vs CUDA graph
for label in ("eager", "graph"):
    if label == "graph":
        eng.capture_graphs()
    for _ in range(2):
        eng.forward(500.0)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        eng.forward(500.0)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    fl = eng.stats("full")[1]
    print(f"{label}: UNet forward n={2 * n_lat}: {ms:.2f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
eps_g = eng.forward(951.0)
eng.x_in.copy_(x)
eps_g = eng.forward(951.0).clone()
torch.cuda.synchronize()
print("graph replay equals eager:", torch.equal(eps_g, eps) if False else (eps_g - ref).abs().max().item())
