"""Data-parallel invariance check (run with torchrun, one rank per GPU): a CD-1 update and a clamped
CD update of a sharded minibatch must equal the single-GPU update of the whole minibatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as td
import multimodal_idbn_b200 as M

rank = M.dist.init_from_env("nccl")
world = td.get_world_size()
dev = torch.device("cuda", torch.cuda.current_device())
M.set_precision(os.environ.get("PREC", "fp32"))
V, H, Dz, B = 532, 256, 500, 64 * world


def make():
    torch.manual_seed(1)
    r = M.RBM(V, H, 0.1, 1e-4, 0.5, dynamic_lr=True, final_momentum=0.95, sparsity=True, sparsity_factor=0.1,
              softmax_groups=[(Dz, V)]).to(dev)
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        r.W.data.copy_(torch.randn(V, H, generator=g) * 0.1)
    r.set_rng(99, 0)
    return r


g = torch.Generator().manual_seed(3)
data = torch.cat([torch.rand(B, Dz, generator=g), torch.nn.functional.one_hot(torch.randint(0, 32, (B,), generator=g), 32).float()], 1).to(dev)
vk = torch.zeros(B, V, device=dev); km = torch.zeros(B, V, device=dev); vk[:, Dz:] = data[:, Dz:]; km[:, Dz:] = 1

ref = make()
l1 = ref.train_epoch(data, 0, 1, CD=2)
l2 = ref.train_epoch_clamped(vk, km, 0, 1, CD=1, cond_init_steps=10, sample_h=False)

M.dist.enable(p2p={"1": True, "0": False}.get(os.environ.get("P2P", ""), None))
lo, hi = M.dist.shard_rows(B, rank, world)
r = make()
d1 = r.train_epoch(data[lo:hi], 0, 1, CD=2)
d2 = r.train_epoch_clamped(vk[lo:hi], km[lo:hi], 0, 1, CD=1, cond_init_steps=10, sample_h=False)
r.sync_momenta()
tol = dict(rtol=1e-4, atol=1e-6) if M.get_precision() == "fp32" else dict(rtol=1e-2, atol=1e-4)
for a, b, n in ((r.W, ref.W, "W"), (r.hid_bias, ref.hid_bias, "hb"), (r.vis_bias, ref.vis_bias, "vb"),
                (r.W_m, ref.W_m, "Wm"), (d1, l1, "loss"), (d2, l2, "loss_clamped")):
    torch.testing.assert_close(a.detach(), b.detach(), msg=lambda m, n=n: f"{n}: {m}", **tol)
# replicas identical
w = r.W.detach().clone(); td.broadcast(w, 0)
assert torch.equal(w, r.W.detach())
td.barrier()
if rank == 0:
    mc = bool(r.__dict__.get("_p2p") and r._p2p["peers"].stats_mc)
    print(f"dp_check ok: world={world} precision={M.get_precision()} p2p={M.dist.state().p2p} multicast={mc}")
td.destroy_process_group()
