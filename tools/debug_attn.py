"""GPU debugging aid for the attention kernel: small cases first, with error localisation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from sonicdiffusionbayeslab_b200 import kernels as k

dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)


def run(B, H, Sq, Sk, d, causal=False):
    C = H * d
    q = torch.randn(B * Sq, C, device=dev, generator=g).bfloat16()
    kk = torch.randn(B * Sk, C, device=dev, generator=g).bfloat16()
    v = torch.randn(B * Sk, C, device=dev, generator=g).bfloat16()
    out = k.attention(q, kk, v, batch=B, heads=H, seq_q=Sq, seq_k=Sk, head_dim=d, causal=causal)
    torch.cuda.synchronize()
    qf = q.float().reshape(B, Sq, H, d).transpose(1, 2)
    kf = kk.float().reshape(B, Sk, H, d).transpose(1, 2)
    vf = v.float().reshape(B, Sk, H, d).transpose(1, 2)
    ref = F.scaled_dot_product_attention(qf, kf, vf, is_causal=causal).transpose(1, 2).reshape(B * Sq, C)
    err = (out.float() - ref).abs()
    print(f"attn B{B} H{H} Sq{Sq} Sk{Sk} d{d} causal={int(causal)}: max_abs={err.max().item():.3e} ref_max={ref.abs().max().item():.3f}",
          flush=True)
    if err.max().item() > 0.05:
        bad = (err > 0.05).nonzero()
        print("  bad", len(bad), "of", err.numel(), "rows", torch.unique(bad[:, 0])[:12].tolist(), "cols",
              torch.unique(bad[:, 1])[:24].tolist())
        print("  out", out[0, :8].float().tolist())
        print("  ref", ref[0, :8].tolist())


for cfg in [(1, 1, 128, 128, 64), (1, 1, 128, 128, 40), (1, 1, 128, 256, 64), (1, 2, 256, 384, 40),
            (1, 1, 128, 77, 64), (1, 1, 64, 64, 160), (1, 1, 128, 128, 80), (1, 1, 128, 256, 160),
            (2, 8, 1024, 1024, 80), (1, 1, 256, 256, 40), (1, 2, 384, 320, 64), (2, 3, 197, 197, 64),
            (1, 2, 300, 77, 40), (2, 8, 4096, 4096, 40), (2, 8, 4096, 77, 40), (1, 1, 256, 1024, 48)]:
    run(*cfg)
for cfg in [(1, 1, 256, 80, 40), (1, 2, 256, 96, 64), (1, 1, 1024, 90, 40), (1, 2, 1152, 100, 64), (1, 1, 1024, 64, 40), (2, 3, 2048, 128, 48), (1, 1, 1280, 77, 16)]:
    run(*cfg)
run(2, 4, 200, 200, 64, causal=True)
run(1, 2, 512, 512, 40, causal=True)
run(2, 8, 77, 77, 64, causal=True)

def bench(B, H, Sq, Sk, d, reps=5):
    C = H * d
    q = torch.randn(B * Sq, C, device=dev, generator=g).bfloat16()
    kv = torch.randn(B * Sk, 2 * C, device=dev, generator=g).bfloat16()
    f = lambda: k.attention(q, kv[:, :C], kv[:, C:], batch=B, heads=H, seq_q=Sq, seq_k=Sk, head_dim=d)
    for _ in range(2):
        f()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        f()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print(f"attention B{B} H{H} Sq{Sq} Sk{Sk} d{d}: {ms:.3f} ms  {4 * B * H * Sq * Sk * d / ms / 1e9:.1f} TFLOP/s", flush=True)


for cfg in [(32, 8, 4096, 4096, 40), (32, 8, 4096, 77, 40), (16, 12, 197, 197, 64), (32, 8, 1024, 1024, 80),
            (32, 8, 256, 256, 160), (32, 8, 1024, 77, 80), (32, 8, 256, 77, 160)]:
    bench(*cfg)
