"""Where does a data-parallel update spend its time?  (torchrun, one rank per GPU; layer-0 shape of C2)"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as td
import multimodal_idbn_b200 as M
from multimodal_idbn_b200 import _lib as L

rank = M.dist.init_from_env("nccl")
world = td.get_world_size()
dev = torch.device("cuda", torch.cuda.current_device())
M.set_precision("tf32")
M.dist.enable()
dp = M.dist.state()
V, H, B = int(os.environ.get("V", 10000)), int(os.environ.get("H", 1500)), 64
r = M.RBM(V, H, 0.1, 1e-4, 0.5).to(dev)
x = [(torch.rand(B, V, device=dev) < 0.1).float() for _ in range(4)]
for i in range(5): r.train_epoch(x[i % 4], 0, 1)
ctx, st = r._ctx()
rs = r._struct(training=True)
p2p = r.__dict__.get("_p2p")
stats = r._stats_buffer(ctx, rs)
upd = r._update_struct(0.1, 0.5, B * world, False)
loss = torch.empty((), device=dev)

def timed(name, fn, n=30):
    for _ in range(3): fn()
    td.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{name:34s} {e0.elapsed_time(e1) / n * 1e3:8.1f} us")

rng = L.RngStruct(1, 0, 0)
timed("cd_stats (3 passes + dS)", lambda: ctx.check(ctx.lib.imdbn_cd_stats(ctx.handle, C.byref(rs), L.ptr(x[0]), B, 1, C.byref(rng), None, L.ptr(stats), st), "s"))
if p2p is not None:
    timed("symm-mem barrier", lambda: p2p["hS"].barrier(channel=0))
    def upd_only():
        p2p["hS"].barrier(channel=0)
        ctx.check(ctx.lib.imdbn_dp_update(ctx.handle, C.byref(rs), C.byref(p2p["peers"]), C.byref(upd), L.ptr(loss), st), "u")
        p2p["hS"].barrier(channel=1)
    timed("barrier + dp_update + barrier", upd_only)
    mcS, mcW = p2p["peers"].stats_mc, p2p["peers"].W_mc
    p2p["peers"].stats_mc, p2p["peers"].W_mc = None, None
    timed("  same, unicast loads/stores", upd_only)
    p2p["peers"].stats_mc, p2p["peers"].W_mc = mcS, mcW
else:
    timed("NCCL all-reduce + apply_update", lambda: (dp.all_reduce(stats), ctx.check(ctx.lib.imdbn_apply_update(ctx.handle, C.byref(rs), L.ptr(stats), C.byref(upd), L.ptr(loss), st), "a")))
if p2p is not None:
    peer = (rank + 1) % world
    n = V * H
    remoteW = p2p["hW"].get_buffer(peer, (n,), torch.float32)
    remoteS = p2p["hS"].get_buffer(peer, (n,), torch.float32)
    localS = stats[:n]
    tmp = torch.empty(n // world, device=dev)
    q = n // world
    timed(f"P2P read  {q * 4 / 1e6:.0f} MB (torch copy)", lambda: tmp.copy_(remoteS[:q]))
    timed(f"P2P write {q * 4 / 1e6:.0f} MB (torch copy)", lambda: remoteS[q:2 * q].copy_(tmp) if world > 1 else None)
    timed(f"local copy {q * 4 / 1e6:.0f} MB", lambda: tmp.copy_(localS[:q]))
timed("forward (up pass, B=128)", lambda: r.forward(torch.cat([x[0], x[1]], 0)))
timed("train_epoch_fwd (whole DP update)", lambda: r.train_epoch_fwd(x[0], 0, 1, next_data=x[1]))
td.barrier()
td.destroy_process_group()
