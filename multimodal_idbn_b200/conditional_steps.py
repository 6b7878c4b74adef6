"""Conditional-step diagnostics of the reference (``imdbn/utils/conditional_steps.py``) on the chain
kernel: the single conditional Gibbs step (``_gibbs_conditional_step``, :15-34) and the convergence
traces ``trace_img2txt_cross`` (:40-126) / ``trace_txt2img_cross`` (:132-238).

The reference traces one sample at a time and reads 6-10 scalars back per step.  Here a whole panel
of samples is stepped together, the per-step measurements stay on the device, and the convergence
rule is evaluated on the host once per trace -- the returned dictionaries hold the same lists,
truncated at the same step, as the reference's.  Plotting / W&B logging (:277-361, 453-471) is out
of scope.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn.functional as F

from . import _lib as L


@torch.no_grad()
def _gibbs_conditional_step(rbm, v, v_known, known_mask, sample_h=False, sample_v=False):
    """One conditional Gibbs step re-clamped to the ORIGINAL ``v_known``; returns
    ``(v_next, v_prob)`` (conditional_steps.py:15-34).  Draws: 0 = U[B,H], 1 = U[B,V], 2 = categorical."""
    return rbm._run_chain(L.CHAIN_COND_GIBBS, v_known, known_mask, 1, sample_h=sample_h,
                          sample_v=sample_v, final_free=False, v_init=v, draw0=-1, want_vprob=True)


@torch.no_grad()
def trace_img2txt_cross(model, img, lbl_onehot=None, max_steps=70, sample_h=False, sample_v=False,
                        eps_l1=1e-3, stable_steps=3, gap_thresh=0.25):
    """IMG->TXT convergence trace with z clamped (conditional_steps.py:40-126).  ``img`` may hold one
    sample (returns a dict, like the reference) or a batch (returns a list of dicts)."""
    dev = model.device
    x = img.view(img.size(0), -1).float().to(dev) if img.dim() > 2 else img.float().to(dev)
    z = model.image_idbn.represent(x)
    B = z.size(0)
    Dz = getattr(model, "Dz_img", z.size(1))
    K = lbl_onehot.size(1) if lbl_onehot is not None else getattr(model, "num_labels", 32)
    V = Dz + K
    jr = model.joint_rbm

    v_known = torch.zeros(B, V, device=dev)
    v_known[:, :Dz] = z
    km = torch.zeros_like(v_known)
    km[:, :Dz] = 1.0
    rng0 = jr._next_rng()
    from .rbm import random_field
    v = v_known * km + (1 - km) * random_field(rng0, 0, B, V, dev)
    y_prev = jr.visible_probs(jr.forward(v))[:, Dz:]
    pred0 = y_prev.argmax(dim=1)
    gt = lbl_onehot.to(dev).argmax(dim=1) if lbl_onehot is not None else None

    rec = {k: [] for k in ("p1", "p2", "k1", "k2", "pgt", "l1", "pred")}
    for _ in range(int(max_steps)):
        v, v_prob = _gibbs_conditional_step(jr, v, v_known, km, sample_h=sample_h, sample_v=sample_v)
        y_soft = v_prob[:, Dz:]
        vals, idxs = y_soft.topk(2, dim=1)
        rec["p1"].append(vals[:, 0]); rec["p2"].append(vals[:, 1])
        rec["k1"].append(idxs[:, 0]); rec["k2"].append(idxs[:, 1])
        if gt is not None:
            rec["pgt"].append(y_soft.gather(1, gt[:, None])[:, 0])
        rec["l1"].append((y_soft - y_prev).abs().sum(dim=1))
        rec["pred"].append(y_soft.argmax(dim=1))
        y_prev = y_soft
    host = {k: (torch.stack(v, 0).cpu() if v else None) for k, v in rec.items()}   # [steps, B]
    pred0 = pred0.cpu()
    gt_h = gt.cpu() if gt is not None else None

    outs = []
    for b in range(B):
        pred_cur, streak, conv, n = int(pred0[b]), 0, max_steps + 1, int(max_steps)
        for t in range(1, int(max_steps) + 1):
            p_new = int(host["pred"][t - 1, b])
            streak = streak + 1 if p_new == pred_cur else 1
            pred_cur = p_new
            gap = float(host["p1"][t - 1, b]) - float(host["p2"][t - 1, b])
            if float(host["l1"][t - 1, b]) < eps_l1 and streak >= stable_steps and gap >= gap_thresh:
                conv, n = t, t
                break
        col = lambda k, cast: [cast(x) for x in host[k][:n, b]]
        p1, p2 = col("p1", float), col("p2", float)
        outs.append({
            "dir": "img2txt", "steps_to_converge": conv, "p_top1": p1, "p_top2": p2,
            "p_gap": [a - c for a, c in zip(p1, p2)],
            "p_gt": col("pgt", float) if gt_h is not None else None,
            "l1": col("l1", float), "predT": pred_cur,
            "top1_idx": col("k1", int), "top2_idx": col("k2", int),
            "gt_idx": int(gt_h[b]) if gt_h is not None else None,
        })
    return outs[0] if B == 1 else outs


@torch.no_grad()
def trace_txt2img_cross(model, img, lbl_onehot, max_steps=70, sample_h=False, sample_v=False,
                        eps_z=1e-3, mse_tol=1e-5, patience=3, ema_beta: float = 0.0):
    """TXT->IMG convergence trace with y clamped, decoding through the image iDBN every step
    (conditional_steps.py:132-238).  One sample -> dict; a batch -> list of dicts."""
    dev = model.device
    img_gt = img.to(dev).view(img.size(0), -1).float()
    B = img_gt.size(0)
    Dz = getattr(model, "Dz_img", int(model.image_idbn.layers[-1].num_hidden))
    K = getattr(model, "num_labels", lbl_onehot.size(1))
    V = Dz + K
    jr = model.joint_rbm
    y = lbl_onehot.to(dev).float()
    v_known = torch.zeros(B, V, device=dev)
    v_known[:, Dz:] = y
    km = torch.zeros_like(v_known)
    km[:, Dz:] = 1.0
    v = v_known.clone()
    if getattr(model, "z_class_mean", None) is not None:
        v[:, :Dz] = model.z_class_mean[y.argmax(dim=1)]
    else:
        v = jr.visible_probs(jr.forward(v_known)) * (1 - km) + v_known * km

    z_prev = v[:, :Dz].clone()
    dz_l, mse_l = [], []
    for _ in range(int(max_steps)):
        v, v_prob = _gibbs_conditional_step(jr, v, v_known, km, sample_h=sample_h, sample_v=sample_v)
        z_soft = v_prob[:, :Dz]
        z_new = (1.0 - ema_beta) * z_prev + ema_beta * z_soft if ema_beta > 0.0 else z_soft
        rec = model.image_idbn.decode(z_new.contiguous()).view_as(img_gt)
        mse_l.append(((rec - img_gt) ** 2).mean(dim=1))
        dz_l.append(torch.norm(z_new - z_prev, p=2, dim=1))
        z_prev = z_new
    dz_h = torch.stack(dz_l, 0).cpu()
    mse_h = torch.stack(mse_l, 0).cpu()

    outs = []
    for b in range(B):
        best, no_imp, conv, n = float("inf"), 0, max_steps + 1, int(max_steps)
        for t in range(1, int(max_steps) + 1):
            mse, dz = float(mse_h[t - 1, b]), float(dz_h[t - 1, b])
            if dz < eps_z:
                if mse + 1e-12 < best - mse_tol:
                    best, no_imp = mse, 0
                else:
                    no_imp += 1
                if no_imp >= patience:
                    conv, n = t, t
                    break
            else:
                if mse + 1e-12 < best - mse_tol:
                    best = mse
                no_imp = 0
        outs.append({"dir": "txt2img", "steps_to_converge": conv,
                     "z_l2": [float(x) for x in dz_h[:n, b]],
                     "image_mse": [float(x) for x in mse_h[:n, b]], "best_mse": best})
    return outs[0] if B == 1 else outs


@torch.no_grad()
def run_cross_panel(model, imgs, lbls, max_steps=70, **kw):
    """Batched counterpart of ``run_and_log_cross_panel`` (conditional_steps.py:474-555) without the
    plotting: both traces for every sample of the panel, stepped together."""
    a = trace_img2txt_cross(model, imgs, lbls, max_steps=max_steps,
                            **{k: v for k, v in kw.items() if k in ("sample_h", "sample_v", "eps_l1",
                                                                    "stable_steps", "gap_thresh")})
    b = trace_txt2img_cross(model, imgs, lbls, max_steps=max_steps,
                            **{k: v for k, v in kw.items() if k in ("sample_h", "sample_v", "eps_z",
                                                                    "mse_tol", "patience", "ema_beta")})
    a = a if isinstance(a, list) else [a]
    b = b if isinstance(b, list) else [b]
    return {"img2txt": a, "txt2img": b,
            "steps_img2txt": [o["steps_to_converge"] for o in a],
            "steps_txt2img": [o["steps_to_converge"] for o in b]}
