"""Host enqueue cost per C2 train step: plain vs pipelined layers (short runs so the launch queue never fills)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
M.set_precision("tf32")
dev = torch.device("cuda")
os.chdir("/tmp")
P = dict(LEARNING_RATE=0.1, WEIGHT_PENALTY=1e-4, INIT_MOMENTUM=0.5, FINAL_MOMENTUM=0.95, LEARNING_RATE_DYNAMIC=True, CD=1)
x = (torch.rand(8, 64, 10000, device=dev) < 0.1).float()
for pipe in (False, True):
    m = M.iDBN([10000, 1500, 500], P, None, None, dev)
    m.pipeline_layers = pipe
    for i in range(20): m.train_step(x[i % 8], 0, 1, next_v=x[(i + 1) % 8])
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(30): m.train_step(x[i % 8], 0, 1, next_v=x[(i + 1) % 8])
        best = min(best, (time.perf_counter() - t0) / 30 * 1e6)
    torch.cuda.synchronize()
    # GPU-side time with a deep queue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(300): m.train_step(x[i % 8], 0, 1, next_v=x[(i + 1) % 8])
    m.sync(); e1.record(); torch.cuda.synchronize()
    print(f"pipeline={pipe}: host enqueue {best:.1f} us/step; steady state {e0.elapsed_time(e1) / 300 * 1e3:.1f} us/step")
import cProfile, pstats
for pipe in (False, True):
    m = M.iDBN([10000, 1500, 500], P, None, None, dev)
    m.pipeline_layers = pipe
    for i in range(60): m.train_step(x[i % 8], 0, 1, next_v=x[(i + 1) % 8])
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for i in range(40): m.train_step(x[i % 8], 0, 1, next_v=x[(i + 1) % 8])
    pr.disable(); m.sync(); torch.cuda.synchronize()
    print("pipeline", pipe)
    pstats.Stats(pr).sort_stats("tottime").print_stats(6)
