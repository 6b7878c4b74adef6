"""One TXT->IMG annealing run on the persistent tensor-core chain kernel (profiling target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
M.set_precision("tf32")
dev = "cuda"
V, H, Dz, K = 532, 256, 500, 32
B = int(sys.argv[1]) if len(sys.argv) > 1 else 14208
torch.manual_seed(0)
r = M.RBM(V, H, 0.04, 1e-4, 0.5, softmax_groups=[(Dz, V)]).to(dev)
y = torch.nn.functional.one_hot(torch.randint(0, K, (B,), device=dev), K).float()
vk = torch.zeros(B, V, device=dev); km = torch.zeros(B, V, device=dev); vk[:, Dz:] = y; km[:, Dz:] = 1
r._mu_pull = {"mu_k": torch.rand(B, Dz, device=dev), "eta0": 0.15}
for _ in range(2):
    out = r.noisy_meanfield_annealed(vk, km, n_steps=50, clamp_suffix=Dz)
torch.cuda.synchronize()
print("ok", float(out.mean()))
