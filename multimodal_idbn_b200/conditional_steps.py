"""Conditional-step diagnostics of the reference (``imdbn/utils/conditional_steps.py``) on the chain
kernel: the single conditional Gibbs step (``_gibbs_conditional_step``, :15-34) and the convergence
traces ``trace_img2txt_cross`` (:40-126) / ``trace_txt2img_cross`` (:132-238).

The reference traces one sample at a time and reads 6-10 scalars back per step.  Here a whole panel
of samples is stepped together, the per-step measurements stay on the device, and the convergence
rule is evaluated on the host once per trace -- the returned dictionaries hold the same lists,
truncated at the same step, as the reference's.  The reference's drivers around the traces are here too
(``pick_fixed_val_case`` :244-275, ``run_and_log_cross_fixed_case`` :364-387, ``build_or_get_fixed_val_panel``
:392-433, ``run_and_log_cross_panel`` :474-555, ``run_and_log_z_mismatch_check`` :557-646) with the panel stepped as
ONE batch; they log the same scalar dictionaries to ``model.wandb_run`` when there is one.  Rendering
(matplotlib figures, ``wandb.Image``; :277-361, 453-471) is out of scope.

Random numbers of the batched drivers: one call = one stream of the joint RBM's random field, sample i = row i.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
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


# ------------------------------------------------------------------------------------------------
# drivers (the reference's W&B entry points, minus the rendering)
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def pick_fixed_val_case(model, target_label: Optional[int] = None, within_batch_index: int = 0):
    """Pick and cache ONE validation sample (conditional_steps.py:244-275)."""
    dev = model.device
    if getattr(model, "_fixed_val_case", None) is not None:
        img_cpu, lbl_cpu = model._fixed_val_case
        return img_cpu.to(dev), lbl_cpu.to(dev)
    if model.val_loader is None:
        raise RuntimeError("model.val_loader is None")
    chosen_img, chosen_lbl = None, None
    if target_label is None:
        for imgs, lbls in model.val_loader:
            chosen_img = imgs[within_batch_index:within_batch_index + 1].cpu()
            chosen_lbl = lbls[within_batch_index:within_batch_index + 1].cpu()
            break
    else:
        for imgs, lbls in model.val_loader:
            idx = (lbls.argmax(dim=1) == target_label).nonzero(as_tuple=True)[0]
            if idx.numel() > 0:
                i0 = int(idx[0])
                chosen_img, chosen_lbl = imgs[i0:i0 + 1].cpu(), lbls[i0:i0 + 1].cpu()
                break
        if chosen_img is None:
            imgs, lbls = next(iter(model.val_loader))
            chosen_img, chosen_lbl = imgs[:1].cpu(), lbls[:1].cpu()
    model._fixed_val_case = (chosen_img, chosen_lbl)
    return chosen_img.to(dev), chosen_lbl.to(dev)


def log_cross_case(model, out_img2txt, out_txt2img, epoch: int, tag: str):
    """conditional_steps.py:277-361 renders two matplotlib figures for W&B: out of scope, nothing is logged."""
    return None


@torch.no_grad()
def run_and_log_cross_fixed_case(model, epoch: int, target_label: Optional[int] = None, within_batch_index: int = 0,
                                 max_steps: int = 70, sample_h: bool = False, sample_v: bool = False,
                                 tag: str = "fixed_cross"):
    """Both directions on the same cached validation sample (conditional_steps.py:364-387)."""
    img, lbl = pick_fixed_val_case(model, target_label=target_label, within_batch_index=within_batch_index)
    a = trace_img2txt_cross(model, img, lbl_onehot=lbl, max_steps=max_steps, sample_h=sample_h, sample_v=sample_v)
    b = trace_txt2img_cross(model, img, lbl_onehot=lbl, max_steps=max_steps, sample_h=sample_h, sample_v=sample_v)
    log_cross_case(model, a, b, epoch=epoch, tag=tag)
    return a, b


@torch.no_grad()
def build_or_get_fixed_val_panel(model, per_class: int = 4):
    """Fixed validation panel with ``per_class`` samples per class, cached on the model
    (conditional_steps.py:392-433)."""
    dev = model.device
    if getattr(model, "_fixed_val_panel", None) is not None:
        imgs_cpu, lbls_cpu = model._fixed_val_panel
        return imgs_cpu.to(dev), lbls_cpu.to(dev)
    if model.val_loader is None:
        raise RuntimeError("val_loader is None")
    K = getattr(model, "num_labels", 32)
    buckets = [[] for _ in range(K)]
    for imgs, lbls in model.val_loader:
        cls = lbls.argmax(dim=1).tolist()                      # one read-back per batch instead of one per sample
        for i, c in enumerate(cls):
            if len(buckets[c]) < per_class:
                buckets[c].append((imgs[i:i + 1].cpu(), lbls[i:i + 1].cpu()))
        if all(len(b) >= per_class for b in buckets):
            break
    imgs_list = [x for b in buckets for (x, _) in b]
    lbls_list = [y for b in buckets for (_, y) in b]
    if not imgs_list:
        imgs, lbls = next(iter(model.val_loader))
        imgs_list, lbls_list = [imgs[:1].cpu()], [lbls[:1].cpu()]
    model._fixed_val_panel = (torch.cat(imgs_list, 0), torch.cat(lbls_list, 0))
    return model._fixed_val_panel[0].to(dev), model._fixed_val_panel[1].to(dev)


def _steps_stats(steps_list, max_steps):
    """Statistics of steps_to_converge over the converged samples (conditional_steps.py:436-450)."""
    arr = np.asarray(steps_list, dtype=np.int32)
    conv_mask = arr <= max_steps
    conv = arr[conv_mask]
    stats = {"n_total": int(arr.size), "n_converged": int(conv.size),
             "frac_converged": float(conv.size / max(1, arr.size)),
             "mean": float(conv.mean()) if conv.size else None,
             "p50": float(np.percentile(conv, 50)) if conv.size else None,
             "p95": float(np.percentile(conv, 95)) if conv.size else None}
    return stats, conv_mask


@torch.no_grad()
def run_and_log_cross_panel(model, epoch: int, per_class: int = 4, max_steps: int = 70, sample_h: bool = False,
                            sample_v: bool = False, tag: str = "panel"):
    """IMG->TXT and TXT->IMG on the fixed panel, aggregated (conditional_steps.py:474-555).  The reference traces
    the up to 128 samples one by one; here each direction is ONE batched trace.  Returns the reference's dictionary
    and logs its ``summary`` to ``model.wandb_run`` (the two histogram figures are out of scope)."""
    imgs, lbls = build_or_get_fixed_val_panel(model, per_class=per_class)
    a = trace_img2txt_cross(model, imgs, lbl_onehot=lbls, max_steps=max_steps, sample_h=sample_h, sample_v=sample_v)
    b = trace_txt2img_cross(model, imgs, lbl_onehot=lbls, max_steps=max_steps, sample_h=sample_h, sample_v=sample_v)
    a = a if isinstance(a, list) else [a]
    b = b if isinstance(b, list) else [b]
    i2t_steps = [int(o["steps_to_converge"]) for o in a]
    t2i_steps = [int(o["steps_to_converge"]) for o in b]
    p1 = [float(o["p_top1"][-1]) for o in a if len(o.get("p_top1", [])) > 0]
    gap = [float(o["p_gap"][-1]) for o in a if len(o.get("p_gap", [])) > 0]
    best = [float(o.get("best_mse", float("inf"))) for o in b]
    i2t_stats, _ = _steps_stats(i2t_steps, max_steps)
    t2i_stats, _ = _steps_stats(t2i_steps, max_steps)
    mean_p1 = float(np.mean(p1)) if p1 else None
    mean_gap = float(np.mean(gap)) if gap else None
    mean_best = float(np.mean(best)) if best else None
    run = getattr(model, "wandb_run", None)
    if run is not None:
        summary = {"img2txt/mean": i2t_stats["mean"], "img2txt/p50": i2t_stats["p50"], "img2txt/p95": i2t_stats["p95"],
                   "img2txt/frac_converged": i2t_stats["frac_converged"],
                   "txt2img/mean": t2i_stats["mean"], "txt2img/p50": t2i_stats["p50"], "txt2img/p95": t2i_stats["p95"],
                   "txt2img/frac_converged": t2i_stats["frac_converged"],
                   "img2txt/p_top1_final_mean": mean_p1, "img2txt/p_gap_final_mean": mean_gap,
                   "txt2img/best_mse_mean": mean_best, "n_total": i2t_stats["n_total"]}
        run.log({f"conv/panel/{tag}/summary": summary, "epoch": epoch})
    return {"img2txt": {"steps": i2t_steps, "stats": i2t_stats, "p1_mean": mean_p1, "gap_mean": mean_gap},
            "txt2img": {"steps": t2i_steps, "stats": t2i_stats, "best_mse_mean": mean_best}}


@torch.no_grad()
def z_mismatch_stats(model, imgs, lbls, max_steps: int = 20, sample_h: bool = False, sample_v: bool = False):
    """The arithmetic of ``run_and_log_z_mismatch_check`` (conditional_steps.py:576-628) on one batch: z from the
    image iDBN against z reached from the clamped label after ``max_steps`` conditional steps (random start of
    the free units), with global statistics and the mean per-sample cosine."""
    dev = model.device
    imgs = imgs.to(dev)
    lbls = lbls.to(dev).float()
    B = imgs.size(0)
    z_img = model.image_idbn.represent(imgs.view(B, -1))
    Dz = z_img.size(1)
    K = getattr(model, "num_labels", lbls.size(1))
    jr = model.joint_rbm
    v_known = torch.zeros(B, Dz + K, device=dev)
    v_known[:, Dz:] = lbls
    km = torch.zeros_like(v_known)
    km[:, Dz:] = 1.0
    from .rbm import random_field
    v = v_known * km + (1 - km) * random_field(jr._next_rng(), 0, B, Dz + K, dev)
    v_prob = v
    for _ in range(int(max_steps)):
        v, v_prob = _gibbs_conditional_step(jr, v, v_known, km, sample_h=sample_h, sample_v=sample_v)
    z_y = v_prob[:, :Dz]

    def _stats(t):
        return {"mean": float(t.mean()), "std": float(t.std(unbiased=False)),
                "q10": float(t.quantile(0.10)), "q90": float(t.quantile(0.90))}

    zi = z_img / (z_img.norm(dim=1, p=2, keepdim=True) + 1e-12)
    zy = z_y / (z_y.norm(dim=1, p=2, keepdim=True) + 1e-12)
    cosine = (zi * zy).sum(dim=1).clamp(-1, 1)
    return {"z_img_stats": _stats(z_img), "z_y_stats": _stats(z_y), "cosine_mean": float(cosine.mean()),
            "z_img": z_img, "z_y": z_y, "cosine": cosine}


@torch.no_grad()
def run_and_log_z_mismatch_check(model, epoch: int, max_steps: int = 20, sample_h: bool = False,
                                 sample_v: bool = False, tag: str = "z_check"):
    """conditional_steps.py:557-646: like the reference, does nothing without a W&B run; logs the three scalar
    entries (the two histogram figures are out of scope).  The per-sample TXT->IMG traces the reference runs first
    (:582-592) only feed a variable it never reads, so they are not repeated here."""
    run = getattr(model, "wandb_run", None)
    if run is None:
        return None
    try:
        imgs, lbls = next(iter(model.val_loader))
    except Exception:
        return None
    st = z_mismatch_stats(model, imgs, lbls, max_steps=max_steps, sample_h=sample_h, sample_v=sample_v)
    run.log({f"zcheck/{tag}/z_img_stats": st["z_img_stats"], "epoch": epoch})
    run.log({f"zcheck/{tag}/z_y_stats": st["z_y_stats"], "epoch": epoch})
    run.log({f"zcheck/{tag}/cosine_mean": st["cosine_mean"], "epoch": epoch})
    return st
