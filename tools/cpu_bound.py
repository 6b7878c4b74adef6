"""Is the training loop bound by host enqueue time or by the GPU?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
M.set_precision("tf32")
dev = torch.device("cuda")
os.chdir("/tmp")
P = dict(LEARNING_RATE=0.1, WEIGHT_PENALTY=1e-4, INIT_MOMENTUM=0.5, FINAL_MOMENTUM=0.95, LEARNING_RATE_DYNAMIC=True, CD=1)
m = M.iDBN([10000, 1500, 500], P, None, None, dev)
x = (torch.rand(8, 64, 10000, device=dev) < 0.1).float()
for fused in (True, False):
    for i in range(20): m.train_step(x[i % 8], 0, 1, next_v=x[(i + 1) % 8] if fused else None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 300
    for i in range(n): m.train_step(x[i % 8], 0, 1, next_v=x[(i + 1) % 8] if fused else None)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"fused={fused}: host enqueue {1e6*(t1-t0)/n:.1f} us/step, total {1e6*(t2-t0)/n:.1f} us/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(100): m.train_step(x[i % 8], 0, 1, next_v=x[(i + 1) % 8])
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
