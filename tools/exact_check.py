"""Diagnostics of the exact tensor-core mode ('tf32x2') against the fp64 truth and the other two modes:
max |dp| of the up / down passes at the C1 shape, flipped samples, statistics error, per-pass timings.
    python tools/exact_check.py            (GPU box)"""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
from multimodal_idbn_b200 import _lib as L

DEV = "cuda"
torch.manual_seed(0)


def timeit(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def main():
    for (V, H, B) in [(10000, 1500, 64), (1500, 500, 64), (532, 256, 64), (10000, 1500, 128)]:
        g = torch.Generator().manual_seed(V + H)
        W = torch.randn(V, H, generator=g) / V ** 0.5
        hb = torch.randn(H, generator=g) * 0.1
        vb = torch.randn(V, generator=g) * 0.1
        vbin = (torch.rand(B, V, generator=g) < 0.1).float()
        vreal = torch.rand(B, V, generator=g)
        hreal = torch.rand(B, H, generator=g)
        hbin = (hreal > 0.5).float()
        u = torch.rand(B, H, generator=g)
        r = M.RBM(V, H, 0.1, 1e-4, 0.5).to(DEV)
        with torch.no_grad():
            r.W.data.copy_(W); r.hid_bias.data.copy_(hb); r.vis_bias.data.copy_(vb)
        W64 = W.double()
        truth = {}
        for name, x in (("vbin", vbin), ("vreal", vreal)):
            truth[name] = torch.sigmoid(x.double() @ W64 + hb.double())
        for name, x in (("hbin", hbin), ("hreal", hreal)):
            truth[name] = torch.sigmoid(x.double() @ W64.T + vb.double())
        ref32 = torch.sigmoid(vbin @ W + hb)
        print(f"--- V={V} H={H} B={B}   (fp32 MKL vs fp64: {float((ref32.double() - truth['vbin']).abs().max()):.2e})")
        for mode in ("fp32", "tf32", "tf32x2"):
            M.set_precision(mode)
            out = {}
            out["vbin"] = r.forward(vbin.to(DEV)).cpu()
            out["vreal"] = r.forward(vreal.to(DEV)).cpu()
            out["hbin"] = r.visible_probs(hbin.to(DEV)).cpu()
            out["hreal"] = r.visible_probs(hreal.to(DEV)).cpu()
            errs = {k: float((out[k].double() - truth[k]).abs().max()) for k in out}
            flips = int(((out["vbin"] > u) != (truth["vbin"].float() > u)).sum())
            xb, xh = vbin.to(DEV), hbin.to(DEV)
            t_up = timeit(lambda: r.forward(xb))
            t_dn = timeit(lambda: r.visible_probs(xh))
            print(f"{mode:7s} max|dp| up(bin) {errs['vbin']:.2e} up(real) {errs['vreal']:.2e} down(bin) {errs['hbin']:.2e} "
                  f"down(real) {errs['hreal']:.2e}  flips {flips}/{B*H}   up {t_up:.1f} us  down {t_dn:.1f} us")
            # statistics
            vp, vn = vbin.to(DEV), (torch.rand(B, V, generator=g) < 0.1).float().to(DEV)
            hp, hn = hreal.to(DEV), torch.rand(B, H, generator=g).to(DEV)
            out_s = torch.empty(V, H, device=DEV)
            ctx, stream = L.context_for(out_s)
            rs = r._struct()
            def stats(a=vp, b=hp, c=vn, d=hn):
                ctx.check(ctx.lib.imdbn_assoc_stats(ctx.handle, C.byref(rs), L.ptr(a), L.ptr(b), L.ptr(c), L.ptr(d), B,
                                                    L.ptr(out_s), stream), "assoc")
            stats()
            ref = (vp.double().T @ hp.double() - vn.double().T @ hn.double()).cpu()
            e1 = float((out_s.cpu().double() - ref).abs().max())
            vr = vreal.to(DEV)
            stats(vr, hp, vr * 0.5, hn)
            ref2 = (vr.double().T @ hp.double() - (vr * 0.5).double().T @ hn.double()).cpu()
            e2 = float((out_s.cpu().double() - ref2).abs().max())
            t_st = timeit(stats)
            print(f"        stats max|err| binary-v {e1:.2e}  real-v {e2:.2e}   (|dS| ~ {float(ref.abs().mean()):.2f})   {t_st:.1f} us")
            # CD-1 update timing
            data = vbin.to(DEV)
            r.set_rng(1, 0)
            t_cd = timeit(lambda: r.train_epoch(data, 0, 1, CD=1), n=20)
            print(f"        train_epoch CD-1: {t_cd:.1f} us")
            with torch.no_grad():
                r.W.data.copy_(W); r.hid_bias.data.copy_(hb); r.vis_bias.data.copy_(vb)
                r.W_m.zero_(); r.hb_m.zero_(); r.vb_m.zero_()
    M.set_precision("fp32")


if __name__ == "__main__":
    main()
