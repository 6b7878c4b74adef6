"""CD-k throughput vs batch size on the widened layer of BASELINE config C5 (10000 -> 4096)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
dev = "cuda"
V, H, K = 10000, 4096, int(os.environ.get("CDK", "1"))
for mode in os.environ.get("MODES", "tf32").split(","):
    M.set_precision(mode)
    r = M.RBM(V, H, 0.1, 1e-4, 0.5).to(dev)
    for B in [64, 256, 1024, 4096, 8192]:
        x = (torch.rand(B, V, device=dev) < 0.1).float()
        for _ in range(2): r.train_epoch(x, 0, 1, CD=K)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n): r.train_epoch(x, 0, 1, CD=K)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        flops = (3 + 2 * K) * 2.0 * B * V * H
        print(f"{mode} B={B:5d} CD-{K}: {ms:9.3f} ms/step  {B/ms*1e3:12.0f} samples/s  {flops/ms/1e9:8.1f} TFLOP/s")
