"""Chain-steps/s of the cross-modal inference kernels (BASELINE config C4 shapes)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
dev = "cuda"
V, H, Dz, K = 532, 256, 500, 32
M.set_precision(os.environ.get("PREC", "tf32"))
r = M.RBM(V, H, 0.04, 1e-4, 0.5, softmax_groups=[(Dz, V)]).to(dev)
with torch.no_grad():
    r.W.data.mul_(3.0)
for B in [64, 4096, 65536]:
    z = torch.rand(B, Dz, device=dev)
    y = torch.nn.functional.one_hot(torch.randint(0, K, (B,), device=dev), K).float()
    vk = torch.zeros(B, V, device=dev); km = torch.zeros(B, V, device=dev); vk[:, :Dz] = z; km[:, :Dz] = 1
    vk2 = torch.zeros(B, V, device=dev); km2 = torch.zeros(B, V, device=dev); vk2[:, Dz:] = y; km2[:, Dz:] = 1
    mu = torch.rand(B, Dz, device=dev)
    for name, fn in [("cond_gibbs(50)", lambda: r.conditional_gibbs(vk, km, n_steps=50, clamp_prefix=Dz)),
                     ("noisy_mf(50)", lambda: r.noisy_meanfield_annealed(vk2, km2, n_steps=50, clamp_suffix=Dz))]:
        r._mu_pull = {"mu_k": mu, "eta0": 0.15} if name.startswith("noisy") else None
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3 if B >= 65536 else 10
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"B={B:6d} {name:16s} {ms:9.3f} ms  {B*50/ms*1e3:.3e} chain-steps/s")
    r._mu_pull = None
