"""Per-kernel time of one CD-1 update at the tensor-bound end (10000 -> 4096, batch 8192, tf32)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
from multimodal_idbn_b200 import _lib as L
M.set_precision("tf32")
dev = "cuda"
V, H, B = 10000, int(os.environ.get("HID", 4096)), int(os.environ.get("B", 8192))
r = M.RBM(V, H, 0.1, 1e-4, 0.5).to(dev)
x = (torch.rand(B, V, device=dev) < 0.1).float()
ctx, st = L.context_for(x)
for _ in range(2): r.train_epoch(x, 0, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): r.train_epoch(x, 0, 1)
e1.record(); torch.cuda.synchronize()
total = e0.elapsed_time(e1) / 3
ctx.profile(True)
for _ in range(3): r.train_epoch(x, 0, 1)
torch.cuda.synchronize()
g = 2.0 * B * V * H / 1e12
acc = 0
for name, kind, mult, n in (("up", 0, 1, 2), ("down", 1, 1, 1), ("stats+update", 2, 2, 1)):
    ms, cnt = ctx.profile_read(kind, V, H)
    ms /= max(1, cnt)
    acc += ms * n
    print(f"{name:14s} {ms:8.3f} ms x{n}  {g * mult / (ms * 1e-3):7.1f} TFLOP/s")
print(f"step {total:.3f} ms; GEMM kernels {acc:.3f} ms; finishes / elementwise {total - acc:.3f} ms; overall {5 * g / (total * 1e-3):.1f} TFLOP/s")
