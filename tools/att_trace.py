"""Timeline of one attention CTA (needs libsonic built with -DSONIC_ATT_TRACE): per sub-tile, when each softmax warp
waited for S, finished its exponentials and signalled P, and when the MMA lane woke / issued."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200 import kernels as k
from sonicdiffusionbayeslab_b200._lib import lib

B, H, S, d = 32, 8, 4096, 40
dev = torch.device("cuda:0")
Cc = H * d
qkv = torch.randn(B * S, 3 * Cc, device=dev).bfloat16()
for _ in range(3):
    k.attention(qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:], batch=B, heads=H, seq_q=S, seq_k=S, head_dim=d)
torch.cuda.synchronize()
buf = (C.c_longlong * (8 * 160 * 4))()
assert lib().sonic_debug_att_trace(buf) == 0
tr = torch.tensor(list(buf)).view(8, 160, 4)
t0 = int(tr[0, 0, 0])
n = 64
print("sub |  MMA: wait_p  woke  pv_issued s_issued | warp2: wait_s got_s exp_done arrived | warp3 ... (cycles since start)")
for t in range(20, 36):
    row = [f"{t:3d} |"] + [f"{int(tr[5, t, i]) - t0:7d}" for i in range(4)]
    for w in (0, 1, 2, 3):
        row.append("|")
        row += [f"{int(tr[w, t, i]) - t0:7d}" for i in range(4)]
    print(" ".join(row))
for w in (0, 1, 2, 3):
    wait = sum(int(tr[w, t, 1] - tr[w, t, 0]) for t in range(n))
    work = sum(int(tr[w, t, 2] - tr[w, t, 1]) for t in range(n))
    sync = sum(int(tr[w, t, 3] - tr[w, t, 2]) for t in range(n))
    print(f"warp {w}: wait_s {wait / n:.0f}  softmax {work / n:.0f}  st-wait+arrive {sync / n:.0f} cycles per sub-tile; "
          f"total {(int(tr[w, n - 1, 3]) - int(tr[w, 0, 0])) / n:.0f}")
mw = sum(int(tr[5, t, 1] - tr[5, t, 0]) for t in range(n)) / n
mi = sum(int(tr[5, t, 3] - tr[5, t, 1]) for t in range(n)) / n
print(f"MMA lane: wait_p {mw:.0f}, issue {mi:.0f} cycles per sub-tile")
lat = sum(int(tr[5, t, 1]) - max(int(tr[w, t, 3]) for w in (0, 1, 2, 3)) for t in range(n)) / n
print(f"p_full arrive(last warp) -> MMA lane awake: {lat:.0f} cycles")
