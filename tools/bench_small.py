"""Time a whole train_epoch_fwd call of a small layer (1500 -> 500, batch 64) with CUDA events around a
run of calls: single-kernel path (csrc/cd_small.cuh) vs the multi-launch path (IMDBN_NO_CD_SMALL=1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
M.set_precision(os.environ.get("PREC", "tf32"))
dev = "cuda"
V, H, B = int(os.environ.get("V", 1500)), int(os.environ.get("H", 500)), 64
r = M.RBM(V, H, 0.1, 1e-4, 0.5).to(dev)
x = [torch.rand(B, V, device=dev) for _ in range(4)]
for i in range(10): r.train_epoch_fwd(x[i % 4], 0, 1, next_data=x[(i + 1) % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 200
e0.record()
for i in range(n): r.train_epoch_fwd(x[i % 4], 0, 1, next_data=x[(i + 1) % 4])
e1.record(); torch.cuda.synchronize()
print(os.environ.get("PREC", "tf32"), "NO_CD_SMALL" if os.environ.get("IMDBN_NO_CD_SMALL") else "cd_small",
      "train_epoch_fwd us/call:", e0.elapsed_time(e1) / n * 1e3)
