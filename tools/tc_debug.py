"""Debug helper for the tensor-core kernels (run on the GPU box)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
from multimodal_idbn_b200 import _lib as L
M.set_precision("tf32")
dev = "cuda"

def assoc(r, vp, hp, vn, hn):
    V, H, B = r.num_visible, r.num_hidden, vp.shape[0]
    out = torch.full((V, H), float("nan"), device=dev)
    ctx, st = L.context_for(out)
    rs = r._struct()
    ctx.check(ctx.lib.imdbn_assoc_stats(ctx.handle, C.byref(rs), L.ptr(vp), L.ptr(hp), L.ptr(vn), L.ptr(hn), B, L.ptr(out), st), "assoc")
    torch.cuda.synchronize()
    return out

for V, H, B in [(128, 128, 32), (128, 128, 8), (256, 256, 64), (532, 256, 64)]:
    r = M.RBM(V, H, 0.1, 0, 0.5).to(dev)
    g = torch.Generator().manual_seed(1)
    vp = (torch.rand(B, V, generator=g) < 0.5).float().to(dev); vn = (torch.rand(B, V, generator=g) < 0.5).float().to(dev)
    hp = (torch.rand(B, H, generator=g) < 0.5).float().to(dev); hn = (torch.rand(B, H, generator=g) < 0.5).float().to(dev)
    z_v, z_h = torch.zeros_like(vp), torch.zeros_like(hp)
    for name, a in [("pos only", (vp, hp, z_v, z_h)), ("neg only", (z_v, z_h, vn, hn)), ("both", (vp, hp, vn, hn))]:
        out = assoc(r, *a)
        ref = a[0].T @ a[1] - a[2].T @ a[3]
        err = (out - ref).abs()
        print(V, H, B, name, "max err", float(err.max()), "nan", int(torch.isnan(out).sum()), "frac bad", float((err > 1e-3).float().mean()))
        if float(err.max()) > 1e-3 and V == 128 and B == 32 and name == "pos only":
            bad = (err > 1e-3)
            print(" bad rows:", bad.any(1).nonzero().flatten()[:40].tolist())
            print(" bad cols:", bad.any(0).nonzero().flatten()[:40].tolist())
            print(" out[0,:8]", out[0, :8].tolist(), "ref", ref[0, :8].tolist())
            # does out equal ref with permuted rows/cols?
            print(" sum out", float(out.nan_to_num().sum()), "sum ref", float(ref.sum()))
