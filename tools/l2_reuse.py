"""Does W (60 MB) survive in the 126 MB L2 between two passes?  Kernel-only time (CUDA events in the library) of the
SAME pass repeated back to back, against the pass inside a CD update (where a 240 MB update streams through L2 first)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
from multimodal_idbn_b200 import _lib as L

DEV = "cuda"
for (V, H) in [(10000, 1500), (10000, 750), (5000, 1500)]:
    torch.manual_seed(0)
    r = M.RBM(V, H, 0.1, 1e-4, 0.5).to(DEV)
    x = (torch.rand(64, V, device=DEV) < 0.1).float()
    h = (torch.rand(64, H, device=DEV) < 0.5).float()
    for mode in ("tf32", "tf32x2"):
        M.set_precision(mode)
        ctx, _ = L.context_for(x)
        for _ in range(3):
            r.forward(x); r.visible_probs(h)
        torch.cuda.synchronize()
        ctx.profile(True)
        for _ in range(20):
            r.forward(x)
        for _ in range(20):
            r.visible_probs(h)
        torch.cuda.synchronize()
        up = ctx.profile_read(L.KERNEL_UP, V, H); dn = ctx.profile_read(L.KERNEL_DOWN, V, H)
        ctx.profile(False)
        print(f"{mode} {V}x{H} ({V*H*4/1e6:.0f} MB): repeated up {up[0]/up[1]*1e3:.1f} us, repeated down {dn[0]/dn[1]*1e3:.1f} us", flush=True)
