"""Timeline of one CTA of the two-tile attention kernel (needs libsonic built with -DSONIC_ATT_TRACE): per (sub-tile,
tile) item, how long each softmax warp waited for S, computed, and synchronised; when the MMA lane woke and issued."""
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
n = 128
print("item | MMA: wait_p woke pv_issued s_issued | warp0: wait_s got_s exp_done arrived | warp1 | warp2 | warp3")
for it in range(40, 56):
    row = [f"{it:3d} |"] + [f"{int(tr[5, it, i]) - t0:7d}" for i in range(4)]
    for w in (0, 1, 2, 3):
        row.append("|")
        row += [f"{int(tr[w, it, i]) - t0:7d}" for i in range(4)]
    print(" ".join(row))
for w in (0, 1, 2, 3):
    its = range(4, n - 2)
    wait = sum(int(tr[w, t, 1] - tr[w, t, 0]) for t in its) / len(its)
    work = sum(int(tr[w, t, 2] - tr[w, t, 1]) for t in its) / len(its)
    sync = sum(int(tr[w, t, 3] - tr[w, t, 2]) for t in its) / len(its)
    print(f"warp {w}: wait_s {wait:.0f}  ld+softmax+st-issue {work:.0f}  st-wait+arrive {sync:.0f} cycles per item; "
          f"period {(int(tr[w, n - 3, 3]) - int(tr[w, 4, 0])) / (n - 7):.0f}")
its = range(4, n - 2)
mw = sum(int(tr[5, t, 1] - tr[5, t, 0]) for t in its) / len(its)
mp = sum(int(tr[5, t, 2] - tr[5, t, 1]) for t in its) / len(its)
ms = sum(int(tr[5, t, 3] - tr[5, t, 2]) for t in its) / len(its)
print(f"MMA lane: wait_p {mw:.0f}, v-wait + PV issue {mp:.0f}, k-wait + S issue {ms:.0f} cycles per item")
lat = sum(int(tr[5, t, 1]) - max(int(tr[w, t, 3]) for w in (0, 1, 2, 3)) for t in its) / len(its)
print(f"p_full arrive(last warp) -> MMA lane awake: {lat:.0f} cycles")
lat2 = sum(min(int(tr[w, t + 2, 1]) for w in (0, 1, 2, 3)) - int(tr[5, t, 3]) for t in range(4, n - 4)) / (n - 8)
print(f"S(t+1) issued -> first softmax warp has it: {lat2:.0f} cycles")
