"""The Gibbs chains inside the reference's iMDBN logging helpers (``imdbn/utils/imdbn_logging.py``), on the chain
kernel.

``log_latent_trajectory_with_recon_panel`` (:23-255), ``log_pca3_trajectory`` (:257-331),
``log_pca3_trajectory_with_recon_panel`` (:333-542) and ``log_vecdb_neighbors_for_traj`` (:703-893) each re-state, in
line, the same label-clamped chain for ONE sample (:124-125 / 303-311 / 465-476 / 767-775):

    v_0 = [z_class_mean[y] | y]                (or one mean-field sweep from [0 | y] without class means)
    h_t ~ Bernoulli(p(h | v_t));  v_{t+1} = p(v | h_t) with the label block re-clamped

and keep ``z_t = v_t[:, :Dz]`` (plus, in the panel variants, ``decode(z_t)``) for a PCA plot.  Here that chain is one
function over a BATCH of labels, stepped by ``imdbn_run_chain`` (``_gibbs_conditional_step`` with sampled hidden
units: ``torch.bernoulli(p)`` is ``p > U``); the PCA / matplotlib / W&B rendering around it is out of scope.
Random numbers: step t of a call uses the joint RBM's stream ``s0 + t``, draw 0, sample i = row i.
"""
from __future__ import annotations

from typing import Optional

import torch

from .conditional_steps import _gibbs_conditional_step


@torch.no_grad()
def label_clamped_trajectory(model, y_onehot: torch.Tensor, steps: Optional[int] = None, decode: bool = False):
    """Trajectory of the image latent under label-clamped Gibbs sampling (imdbn_logging.py:289-311 and its three
    copies).  ``y_onehot`` [B, K].  Returns ``z_traj`` [steps + 1, B, Dz] on the device and, with ``decode``,
    ``imgs`` [steps + 1, B, D] = ``image_idbn.decode(z_t)`` (the reconstructions of the panel variants, :462-476)."""
    dev = model.device
    jr = model.joint_rbm
    y = y_onehot.to(dev).float()
    B, K = y.shape
    Dz = int(getattr(model, "Dz_img", jr.num_visible - K))
    V = Dz + K
    T = int(model.cross_steps if steps is None else steps)
    v_known = torch.zeros(B, V, device=dev)
    v_known[:, Dz:] = y
    km = torch.zeros_like(v_known)
    km[:, Dz:] = 1.0
    if getattr(model, "z_class_mean", None) is not None:
        v = v_known.clone()
        v[:, :Dz] = model.z_class_mean[y.argmax(dim=1)]
    else:
        v = jr.visible_probs(jr.forward(v_known)) * (1 - km) + v_known * km
    zs = [v[:, :Dz].clone()]
    imgs = [model.image_idbn.decode(zs[0].contiguous())] if decode else None
    for _ in range(T):
        v, _ = _gibbs_conditional_step(jr, v, v_known, km, sample_h=True, sample_v=False)
        zs.append(v[:, :Dz].clone())
        if decode:
            imgs.append(model.image_idbn.decode(zs[-1].contiguous()))
    z_traj = torch.stack(zs, 0)
    return (z_traj, torch.stack(imgs, 0)) if decode else z_traj
