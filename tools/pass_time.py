"""Kernel-only timings (CUDA events inside the library) of the up / down passes and the statistics + update kernel at the
C1 shape, per precision mode.   python tools/pass_time.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
from multimodal_idbn_b200 import _lib as L

DEV = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
V, H = 10000, 1500
torch.manual_seed(0)
r = M.RBM(V, H, 0.1, 1e-4, 0.5).to(DEV)
x = (torch.rand(B, V, device=DEV) < 0.1).float()
for mode in ("tf32", "tf32x2"):
    M.set_precision(mode)
    ctx, _ = L.context_for(x)
    for _ in range(5):
        r.train_epoch(x, 0, 1, CD=1)
    torch.cuda.synchronize()
    ctx.profile(True)
    for _ in range(20):
        r.train_epoch(x, 0, 1, CD=1)
    torch.cuda.synchronize()
    out = {}
    for name, kind in (("up", L.KERNEL_UP), ("down", L.KERNEL_DOWN), ("stats", L.KERNEL_STATS)):
        tot, cnt = ctx.profile_read(kind, V, H)
        out[name] = round(tot / max(1, cnt) * 1e3, 1)
    ctx.profile(False)
    print(mode, "B", B, out, flush=True)
