"""Drop-in ``imdbn.utils.energy_utils`` (reference utils/energy_utils.py): free energies and the deterministic
IMG->TXT convergence trace, batched on the CUDA kernels.

* ``rbm_free_energy`` (:18-28)                -> ``imdbn_free_energy``
* ``class_free_energies`` (:32-54)            -> ``imdbn_class_free_energies`` (one launch for all B x K candidates)
* ``_deterministic_img2txt_step`` (:61-90)    -> ``imdbn_trace_img2txt`` with one step
* ``trace_single_img2txt`` (:96-196)          -> all steps run in ONE kernel that records the label distribution
  of every step; the trace metrics are derived on the host from a single device-to-host copy (the reference
  issues ~8 ``.item()`` per step).  ``trace_batch_img2txt`` does the same for a whole batch of cases.

The W&B renderers (``log_single_case_energy``, ``run_and_log_fixed_case``, :253-387) are out of scope.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch

from . import _lib as L
from .rbm import rbm_free_energy  # noqa: F401  (re-exported)


def _latents(joint_rbm, z: torch.Tensor, Dz: int) -> torch.Tensor:
    z = L.f32c(z, joint_rbm.W.device)
    if z.dim() != 2 or z.shape[1] != Dz:
        raise ValueError(f"expected latents of shape [B, {Dz}], got {tuple(z.shape)}")
    if not z.is_cuda:
        raise RuntimeError("multimodal_idbn_b200 is CUDA only (no CPU fallback)")
    return z


@torch.no_grad()
def class_free_energies(joint_rbm, z_img_top: torch.Tensor, K: int, Dz: int) -> torch.Tensor:
    """F_k(z) = F([z, e_k]) for k = 1..K; ``[B, Dz] -> [B, K]`` (energy_utils.py:32-54)."""
    if Dz + K != joint_rbm.num_visible:
        raise ValueError("Dz + K must equal the number of visible units of the joint RBM")
    z = _latents(joint_rbm, z_img_top, Dz)
    ctx, st = joint_rbm._ctx()
    out = torch.empty(z.shape[0], K, device=z.device, dtype=torch.float32)
    rs = joint_rbm._struct()
    ctx.check(ctx.lib.imdbn_class_free_energies(ctx.handle, C.byref(rs), L.ptr(z), z.shape[0], int(Dz), L.ptr(out), st),
              "imdbn_class_free_energies")
    return out


def _label_trajectory(joint_rbm, z: torch.Tensor, Dz: int, K: int, steps: int,
                      y_init: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``[steps, B, K]`` label distributions of the mean-field-lite chain with z clamped."""
    if Dz + K != joint_rbm.num_visible:
        raise ValueError("Dz + K must equal the number of visible units of the joint RBM")
    z = _latents(joint_rbm, z, Dz)
    ctx, st = joint_rbm._ctx()
    B = z.shape[0]
    y0 = None if y_init is None else L.f32c(y_init, z.device)
    traj = torch.empty(int(steps), B, K, device=z.device, dtype=torch.float32)
    rs = joint_rbm._struct()
    ctx.check(ctx.lib.imdbn_trace_img2txt(ctx.handle, C.byref(rs), L.ptr(z), B, int(Dz), L.ptr(y0), int(steps),
                                          L.ptr(traj), st), "imdbn_trace_img2txt")
    return traj


@torch.no_grad()
def _deterministic_img2txt_step(joint_rbm, v: torch.Tensor, Dz: int, K: int, softmax_y: bool = True,
                                sample_h: bool = False, sample_v: bool = False) -> torch.Tensor:
    """One mean-field-lite step: v -> h_prob -> v_prob, z re-clamped, y = softmax(sigmoid part) (:61-90).
    Only the deterministic configuration the reference uses (:138) is implemented."""
    if not softmax_y or sample_h or sample_v:
        raise NotImplementedError("only softmax_y=True, sample_h=False, sample_v=False (energy_utils.py:138)")
    v = L.f32c(v, joint_rbm.W.device)
    y = _label_trajectory(joint_rbm, v[:, :Dz].contiguous(), Dz, K, 1, v[:, Dz:Dz + K].contiguous())[0]
    return torch.cat([v[:, :Dz], y], dim=1)


def _trace_metrics(Fk: np.ndarray, traj: np.ndarray, gt: Optional[int], eps_l1: float, stable_steps: int,
                   gap_thresh: float) -> dict:
    """energy_utils.py:126-196 for ONE case from its class energies [K] and label trajectory [T, K]."""
    K = Fk.shape[0]
    steps = traj.shape[0]
    kstar = int(np.argmin(Fk))
    Fmin = Fk[kstar]
    srt = np.sort(Fk)
    margin_energy = float(srt[1] - srt[0])
    p_top1, p_top2, p_gap, p_gt, dF = [], [], [], [], []
    y_prev = np.full(K, np.float32(1.0) / np.float32(K), dtype=np.float32)
    pred_cur, streak, steps_to_conv = 0, 0, steps + 1
    for t in range(1, steps + 1):
        y = traj[t - 1]
        top = np.sort(y)[::-1]
        p1, p2 = float(top[0]), float(top[1])
        gap = p1 - p2
        p_top1.append(p1); p_top2.append(p2); p_gap.append(gap)
        if gt is not None:
            p_gt.append(float(y[gt]))
        pred_new = int(np.argmax(y))
        streak = streak + 1 if pred_new == pred_cur else 1
        pred_cur = pred_new
        dF.append(float(Fk[pred_cur] - Fmin))
        l1 = float(np.abs(y - y_prev).sum(dtype=np.float32))
        if l1 < eps_l1 and streak >= stable_steps and (pred_cur == kstar or gap >= gap_thresh):
            steps_to_conv = t
            break
        y_prev = y
    e = np.exp(-(Fk - Fk.min()).astype(np.float32))
    fe = e / e.sum(dtype=np.float32)
    fe_sorted = np.sort(fe)[::-1]
    return {
        "deltaF_pred_traj": dF, "deltaF_pred_final": dF[-1] if dF else None,
        "p_top1": p_top1, "p_top2": p_top2, "p_gap": p_gap, "p_gt": p_gt if gt is not None else None,
        "p_top1_final": p_top1[-1] if p_top1 else float(1.0 / K), "p_gap_final": p_gap[-1] if p_gap else 0.0,
        "fe_top1_final": float(fe_sorted[0]), "fe_gap_final": float(fe_sorted[0] - fe_sorted[1]),
        "steps_to_converge": steps_to_conv, "kstar": kstar, "predT": pred_cur, "margin_energy": margin_energy,
        "gt": gt,
    }


@torch.no_grad()
def trace_batch_img2txt(model, imgs: torch.Tensor, lbls_onehot: Optional[torch.Tensor], steps: int = 30,
                        eps_l1: float = 1e-3, stable_steps: int = 3, gap_thresh: float = 0.25) -> List[dict]:
    """``trace_single_img2txt`` for every row of a batch: one represent, one class-energy launch, one trace
    launch, one device-to-host copy."""
    dev = model.device
    x = imgs.view(imgs.size(0), -1).float().to(dev) if imgs.dim() > 2 else imgs.float().to(dev)
    z = model.image_idbn.represent(x).clamp(1e-6, 1 - 1e-6)                          # :113
    Dz = getattr(model, "Dz_img", z.size(1))
    K = getattr(model, "num_labels", lbls_onehot.size(1) if lbls_onehot is not None else 32)
    Fk = class_free_energies(model.joint_rbm, z, K, Dz)
    traj = _label_trajectory(model.joint_rbm, z, Dz, K, steps)
    Fk_h, traj_h = Fk.cpu().numpy(), traj.cpu().numpy()
    gts = lbls_onehot.argmax(dim=1).cpu().numpy() if lbls_onehot is not None else None
    return [_trace_metrics(Fk_h[b], traj_h[:, b], int(gts[b]) if gts is not None else None, eps_l1, stable_steps,
                           gap_thresh) for b in range(z.size(0))]


@torch.no_grad()
def trace_single_img2txt(model, img: torch.Tensor, lbl_onehot: Optional[torch.Tensor], steps: int = 30,
                         eps_l1: float = 1e-3, stable_steps: int = 3, gap_thresh: float = 0.25) -> dict:
    """energy_utils.py:96-196 (one case)."""
    return trace_batch_img2txt(model, img[:1], None if lbl_onehot is None else lbl_onehot[:1], steps, eps_l1,
                               stable_steps, gap_thresh)[0]


@torch.no_grad()
def pick_fixed_val_case(model, target_label: Optional[int] = None, within_batch_index: int = 0):
    """A validation sample cached on the model so every epoch looks at the same case (:203-237)."""
    dev = model.device
    if getattr(model, "_fixed_val_case", None) is not None:
        img_cpu, lbl_cpu = model._fixed_val_case
        return img_cpu.to(dev), lbl_cpu.to(dev)
    if model.val_loader is None:
        raise RuntimeError("model.val_loader is None")
    chosen_img = chosen_lbl = None
    if target_label is None:
        for imgs, lbls in model.val_loader:
            chosen_img = imgs[within_batch_index:within_batch_index + 1].cpu()
            chosen_lbl = lbls[within_batch_index:within_batch_index + 1].cpu()
            break
    else:
        for imgs, lbls in model.val_loader:
            idx = (lbls.argmax(dim=1) == target_label).nonzero(as_tuple=True)[0]
            if idx.numel() > 0:
                i0 = int(idx[0].item())
                chosen_img, chosen_lbl = imgs[i0:i0 + 1].cpu(), lbls[i0:i0 + 1].cpu()
                break
        if chosen_img is None:
            imgs, lbls = next(iter(model.val_loader))
            chosen_img, chosen_lbl = imgs[:1].cpu(), lbls[:1].cpu()
    model._fixed_val_case = (chosen_img, chosen_lbl)
    return chosen_img.to(dev), chosen_lbl.to(dev)


@torch.no_grad()
def pick_val_case(model, target_label: Optional[int] = None, batch_idx: int = 0, within_batch_index: int = 0):
    """Backward-compatible alias (:241-246)."""
    return pick_fixed_val_case(model, target_label=target_label, within_batch_index=within_batch_index)
