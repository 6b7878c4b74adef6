"""Time the upper layer's update (1500 -> 500, batch 64) on the small SM partition vs the whole chip."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
from multimodal_idbn_b200 import _lib as L
M.set_precision("tf32")
dev = torch.device("cuda", 0)
V, H, B = 1500, 500, 64
x = [torch.rand(B, V, device=dev) for _ in range(4)]
part = L.sm_partition(0, int(os.environ.get("RESERVE", 16)))
print("partition", part)
def run(stream, limit, name):
    r = M.RBM(V, H, 0.1, 1e-4, 0.5).to(dev)
    with torch.cuda.stream(stream):
        ctx, _ = L.context_for(x[0])
        ctx.set_sm_limit(limit)
        for i in range(20): r.train_epoch_fwd(x[i % 4], 0, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n = 300
        for i in range(n): r.train_epoch_fwd(x[i % 4], 0, 1)
        e1.record(stream); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / n * 1e3:.1f} us per train_epoch_fwd")
run(torch.cuda.current_stream(), 0, "default stream, 148 SMs")
if part:
    run(torch.cuda.ExternalStream(part[1], device=dev), part[3], f"small partition ({part[3]} SMs)")
    run(torch.cuda.ExternalStream(part[0], device=dev), part[2], f"big partition ({part[2]} SMs)")
