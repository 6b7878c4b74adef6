"""Host-side enqueue cost of one C2 train step and of the e2e extras (short runs: the launch queue never fills)."""
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
host = (torch.rand(8, 64, 10000) < 0.1).float().pin_memory()
for i in range(20): m.train_step(x[i % 8], 0, 1, next_v=x[(i + 1) % 8])
torch.cuda.synchronize()
def t(fn, n=40, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n): fn(i)
        best = min(best, (time.perf_counter() - t0) / n * 1e6)
        torch.cuda.synchronize()
    return best
print("train_step host us:", t(lambda i: m.train_step(x[i % 8], 0, 1, next_v=x[(i + 1) % 8])))
cs = torch.cuda.Stream(); main = torch.cuda.current_stream()
def stage(i):
    with torch.cuda.stream(cs):
        b = host[i % 8].to(dev, non_blocking=True); ev = torch.cuda.Event(); ev.record(cs)
    main.wait_event(ev); return b
print("stage host us:", t(stage))
losses = m.train_step(x[0], 0, 1, next_v=x[1]); lh = torch.empty(64, 2).pin_memory()
print("stack+copy host us:", t(lambda i: lh[i % 64].copy_(torch.stack(losses), non_blocking=True)))
print("index host us:", t(lambda i: (x[i % 8], x[(i + 1) % 8])))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(40): m.train_step(x[i % 8], 0, 1, next_v=x[(i + 1) % 8])
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
