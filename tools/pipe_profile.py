"""Per-kernel-kind times of both layers while they run concurrently (pipelined C2 step)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
M.set_precision("tf32")
dev = torch.device("cuda")
os.chdir("/tmp")
P = dict(LEARNING_RATE=0.1, WEIGHT_PENALTY=1e-4, INIT_MOMENTUM=0.5, FINAL_MOMENTUM=0.95, LEARNING_RATE_DYNAMIC=True, CD=1)
m = M.iDBN([10000, 1500, 500], P, None, None, dev)
m.pipeline_layers = True
x = (torch.rand(16, 64, 10000, device=dev) < 0.1).float()
for i in range(800): m.train_step(x[i % 16], 0, 1, next_v=x[(i + 1) % 16])
m.sync(); torch.cuda.synchronize()
st = m._fused
for c in (st["ctx0"], st["ctx1"]): c.profile(True)
for i in range(100): m.train_step(x[i % 16], 0, 1, next_v=x[(i + 1) % 16])
m.sync(); torch.cuda.synchronize()
for name, ctx, (V, H) in (("layer 0 (big partition)", st["ctx0"], (10000, 1500)), ("layer 1 (small partition)", st["ctx1"], (1500, 500))):
    parts = []
    for kn, kind in (("up", 0), ("down", 1), ("stats", 2)):
        ms, n = ctx.profile_read(kind, V, H)
        parts.append(f"{kn} {ms / max(1, n) * 1e3:.1f} us x{n // 100}")
    print(name, "|", " | ".join(parts))
