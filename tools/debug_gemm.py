"""GPU debugging aid: run a few GEMM/conv cases and print where mismatches are."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from sonicdiffusionbayeslab_b200 import kernels as k

torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0))


def report(name, out, ref):
    torch.cuda.synchronize()
    err = (out.float() - ref).abs()
    rel = err.max().item() / ref.abs().max().item()
    print(f"{name}: rel_err={rel:.3e} max_abs={err.max().item():.3e} ref_max={ref.abs().max().item():.3f}")
    if rel > 1e-2:
        bad = (err > 0.05 * ref.abs().max()).nonzero()
        print("   bad count", len(bad), "of", err.numel(), "first", bad[:8].tolist())
        rows = torch.unique(bad[:, 0])
        cols = torch.unique(bad[:, 1])
        print("   bad rows", rows[:16].tolist(), "...", len(rows), " bad cols", cols[:16].tolist(), "...", len(cols))
        print("   out[0,:8]", out[0, :8].float().tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())


g = torch.Generator(device="cuda").manual_seed(0)
for (M, N, K, bn) in [(128, 64, 64, 64), (128, 64, 128, 64), (128, 128, 256, 128), (256, 256, 64, 256), (384, 320, 320, 0)]:
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) / K ** 0.5).bfloat16()
    out = k.conv_gemm(a, w, N, block_n=bn)
    report(f"linear M{M} N{N} K{K} bn{bn}", out, a.float() @ w.float().t())

for (B, H, W, Ci, Co) in [(1, 8, 8, 64, 64), (2, 16, 16, 64, 64), (2, 64, 64, 128, 128)]:
    x = torch.randn(B, H, W, Ci, device=dev, generator=g).bfloat16()
    w = (torch.randn(Co, Ci, 3, 3, device=dev, generator=g) / (9 * Ci) ** 0.5).bfloat16()
    out = k.conv_gemm(x, k.pack_conv3x3_weight(w), Co, taps=9, n_img=B, H=H, W=W)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1).permute(0, 2, 3, 1).reshape(-1, Co)
    report(f"conv B{B} {H}x{W} {Ci}->{Co}", out, ref)

# timing of a big GEMM
M, N, K = 32 * 4096, 320, 2880 // 9
x = torch.randn(32, 64, 64, 320, device=dev, generator=g).bfloat16()
w = (torch.randn(320, 320, 3, 3, device=dev, generator=g) / 54).bfloat16()
wp = k.pack_conv3x3_weight(w)
out = torch.empty(32 * 4096, 320, device=dev, dtype=torch.bfloat16)
for bn in (0, 160, 64):
    for _ in range(3):
        k.conv_gemm(x, wp, 320, taps=9, n_img=32, H=64, W=64, out=out, block_n=bn)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        k.conv_gemm(x, wp, 320, taps=9, n_img=32, H=64, W=64, out=out, block_n=bn)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    fl = 2 * 32 * 4096 * 320 * 2880
    print(f"conv3x3 320->320 @64x64 b32 bn={bn}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
x = torch.randn(32 * 256, 1280, device=dev, generator=g).bfloat16()
w = (torch.randn(1280, 1280, device=dev, generator=g) / 36).bfloat16()
for bn in (0, 256, 128):
    for _ in range(3):
        k.conv_gemm(x, w, 1280, block_n=bn)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        k.conv_gemm(x, w, 1280, block_n=bn)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"linear 8192x1280x1280 bn={bn}: {ms:.3f} ms {2 * 8192 * 1280 * 1280 / ms / 1e9:.1f} TFLOP/s")
