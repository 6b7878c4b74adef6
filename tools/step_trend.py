"""Per-chunk step time over a long run (is there a slow transient?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
M.set_precision("tf32")
dev = torch.device("cuda")
os.chdir("/tmp")
P = dict(LEARNING_RATE=0.1, WEIGHT_PENALTY=1e-4, INIT_MOMENTUM=0.5, FINAL_MOMENTUM=0.95, LEARNING_RATE_DYNAMIC=True, CD=1)
m = M.iDBN([10000, 1500, 500], P, None, None, dev)
m.pipeline_layers = os.environ.get("PIPE", "1") == "1"
m.pipeline_reserve_sms = int(os.environ.get("RESERVE", 16))
NB = int(os.environ.get("NB", 64))
x = (torch.rand(NB, 64, 10000, device=dev) < 0.1).float()
for i in range(NB): m.train_step(x[i % NB], 0, 1, next_v=x[(i + 1) % NB])
m.sync(); torch.cuda.synchronize()
NC = int(os.environ.get('CHUNKS', 12))
evs = [torch.cuda.Event(enable_timing=True) for _ in range(NC + 1)]
import time
k = 0
host = []
evs[0].record()
for c in range(NC):
    t0 = time.perf_counter()
    for i in range(100):
        m.train_step(x[k % NB], 0, 1, next_v=x[(k + 1) % NB]); k += 1
    host.append(round((time.perf_counter() - t0) * 1e4, 1))
    evs[c + 1].record()
m.sync(); torch.cuda.synchronize()
print("host us/step per chunk:       ", host)
print("us/step per 100-step chunk:", [round(evs[c].elapsed_time(evs[c + 1]) * 10, 1) for c in range(NC)])
