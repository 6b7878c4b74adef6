import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
M.set_precision("tf32")
dev = "cuda"
V, H, Dz, K, B = 532, 256, 500, 32, int(os.environ.get("B", "65536"))
r = M.RBM(V, H, 0.04, 1e-4, 0.5, softmax_groups=[(Dz, V)]).to(dev)
z = torch.rand(B, Dz, device=dev)
vk = torch.zeros(B, V, device=dev); km = torch.zeros(B, V, device=dev); vk[:, :Dz] = z; km[:, :Dz] = 1
for _ in range(2):
    out = r.conditional_gibbs(vk, km, n_steps=3)
torch.cuda.synchronize()
print("ok", float(out.sum()))
