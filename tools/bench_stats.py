"""Time the layer-0 statistics+update kernel alone (debug)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
from multimodal_idbn_b200 import _lib as L
M.set_precision(os.environ.get("PREC", "tf32"))
dev = "cuda"
V, H, B = 10000, int(os.environ.get("HID", "1500")), 64
r = M.RBM(V, H, 0.1, 1e-4, 0.5).to(dev)
x = (torch.rand(B, V, device=dev) < 0.1).float()
ctx, st = L.context_for(x)
for _ in range(5): r.train_epoch(x, 0, 1)
ctx.profile(True)
for _ in range(30): r.train_epoch(x, 0, 1)
torch.cuda.synchronize()
for name, kind in (("up", 0), ("down", 1), ("stats", 2)):
    ms, n = ctx.profile_read(kind, V, H)
    print(name, "avg us", ms / max(1, n) * 1e3, "n", n)
