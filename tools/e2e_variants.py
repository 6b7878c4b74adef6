"""Where the end-to-end loop loses time against the device-resident one: the same C2 loop with (a) host batches + losses
into mapped host memory (bench.py's e2e), (b) host batches, losses on the device, (c) resident batches, losses into mapped
host memory, (d) resident batches, losses on the device."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import multimodal_idbn_b200 as M
dev = torch.device("cuda", 0)
M.load_library(); M.set_precision("tf32x2")
os.chdir("/tmp")
g = torch.Generator().manual_seed(1)
host = (torch.rand(64, 64, 10000, generator=g) < 0.1).float().pin_memory()
res = host.to(dev)
model, batch = bench.make_c2(M, dev, "tf32x2", 5, 16, res, seed=0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
loss_host = torch.zeros(N + 8, 2).pin_memory()
def loop(host_in, host_loss):
    src = [(host[i % 64],) for i in range(N)] if host_in else [(res[i % 64],) for i in range(N)]
    it = M.prefetch_to_device(src, dev) if host_in else iter(src)
    cur, i = None, 0
    for b in it:
        nxt = b[0]
        if cur is not None:
            model.train_step(cur, 0, 1, next_v=nxt, loss_out=loss_host[i] if host_loss else None); i += 1
        cur = nxt
    model.train_step(cur, 0, 1, loss_out=loss_host[i] if host_loss else None)
    model.sync()
for name, a, b in (("host batches, host losses", 1, 1), ("host batches, device losses", 1, 0), ("resident, host losses", 0, 1), ("resident, device losses", 0, 0)) * 2:
    torch.cuda.synchronize(); t0 = time.perf_counter(); loop(a, b); torch.cuda.synchronize()
    print(f"{name:32s} {(time.perf_counter() - t0) / N * 1e6:7.1f} us/step", flush=True)
