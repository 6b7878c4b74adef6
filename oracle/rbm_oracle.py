"""CPU oracle: a functional restatement of the reference's RBM / iDBN / iMDBN hot path.

TEST INFRASTRUCTURE ONLY -- the product (``multimodal_idbn_b200``) never imports this file; it is
used by ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` as the checker
and as the "port" CPU baseline.

Every function cites the reference lines it restates (paths relative to the reference tree).  The
arithmetic is torch-CPU fp32 in the reference's operation order, so on one machine the oracle and
the reference agree bit for bit when they are fed the same random numbers (that is how the oracle
is pinned: ``tests/golden/make_golden.py`` runs the real reference with an RNG-injection proxy,
and ``tests/test_oracle_golden.py`` compares).  Pass ``dtype=torch.float64`` states to get a
higher-precision "truth" for tolerance tests.

Random numbers come from ``oracle.philox.RandomField`` (one field per API call).  The draw index
of every tensor is stated in each docstring and is the contract the CUDA kernels follow.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field as dc_field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .philox import RandomField


# --------------------------------------------------------------------------------------------
# state
# --------------------------------------------------------------------------------------------
@dataclass
class RBMState:
    """Parameters + momenta + hyper-parameters of one RBM (imdbn/models/rbm.py:41-79)."""
    W: torch.Tensor          # [V, H]
    hb: torch.Tensor         # [H]
    vb: torch.Tensor         # [V]
    Wm: torch.Tensor
    hbm: torch.Tensor
    vbm: torch.Tensor
    lr: float = 0.1
    weight_decay: float = 1e-4
    momentum: float = 0.5
    final_momentum: float = 0.95
    dynamic_lr: bool = True
    sparsity: bool = False
    sparsity_factor: float = 0.05
    groups: List[Tuple[int, int]] = dc_field(default_factory=list)

    @property
    def V(self) -> int:
        return self.W.shape[0]

    @property
    def H(self) -> int:
        return self.W.shape[1]

    def clone(self) -> "RBMState":
        return RBMState(self.W.clone(), self.hb.clone(), self.vb.clone(), self.Wm.clone(),
                        self.hbm.clone(), self.vbm.clone(), self.lr, self.weight_decay,
                        self.momentum, self.final_momentum, self.dynamic_lr, self.sparsity,
                        self.sparsity_factor, list(self.groups))

    def to(self, dtype) -> "RBMState":
        c = self.clone()
        for n in ("W", "hb", "vb", "Wm", "hbm", "vbm"):
            setattr(c, n, getattr(c, n).to(dtype))
        return c


def new_state(V: int, H: int, seed: int = 0, groups=None, dtype=torch.float32, **hyper) -> RBMState:
    """Reference init: W = randn(V,H)/sqrt(max(1,V)), zero biases and momenta (rbm.py:69-79)."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    W = (torch.randn(V, H, generator=g, dtype=torch.float32) / math.sqrt(max(1, V))).to(dtype)
    z = lambda n: torch.zeros(n, dtype=dtype)
    return RBMState(W, z(H), z(V), torch.zeros_like(W), z(H), z(V),
                    groups=list(groups or []), **hyper)


def _t(a: np.ndarray, like: torch.Tensor) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).to(like.dtype)


def lr_and_momentum(st: RBMState, epoch: int) -> Tuple[float, float]:
    """rbm.py:194-195 / 438-439."""
    lr = st.lr / (1 + 0.01 * epoch) if st.dynamic_lr else st.lr
    mom = st.momentum if epoch <= 5 else st.final_momentum
    return lr, mom


# --------------------------------------------------------------------------------------------
# up / down passes and sampling
# --------------------------------------------------------------------------------------------
def hand_sigmoid(x: torch.Tensor) -> torch.Tensor:
    """rbm.py:19-21 -- the reference's own 1/(1+exp(-x)) (used by ``forward`` only)."""
    return 1 / (1 + torch.exp(-x))


def hidden_probs(st: RBMState, v: torch.Tensor, T: float = 1.0) -> torch.Tensor:
    """p(h|v), rbm.py:81-92."""
    return hand_sigmoid((v @ st.W + st.hb) / max(1e-6, T))


def visible_logits(st: RBMState, h: torch.Tensor, T: float = 1.0) -> torch.Tensor:
    """rbm.py:94-96."""
    return (h @ st.W.T + st.vb) / max(1e-6, T)


def _group_softmax_(p: torch.Tensor, logits: torch.Tensor, groups) -> torch.Tensor:
    for s, e in groups:
        p[:, s:e] = torch.softmax(logits[:, s:e], dim=1)
    return p


def visible_probs(st: RBMState, h: torch.Tensor, T: float = 1.0) -> torch.Tensor:
    """p(v|h) with softmax groups, rbm.py:98-116 (library sigmoid here, not the hand-rolled one)."""
    logits = visible_logits(st, h, T)
    return _group_softmax_(torch.sigmoid(logits), logits, st.groups)


def categorical_index(q: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """Parity definition of ``Categorical(probs=q).sample()`` (rbm.py:130-131).

    ``q`` [B, K] already clamped to [1e-8, 1]; normalised as torch.distributions does
    (q / q.sum(-1)); inverse CDF with a sequential cumulative sum:
        idx = min(K-1, #{j : cdf_j <= u}).
    """
    qn = q / q.sum(dim=-1, keepdim=True)
    cdf = torch.cumsum(qn, dim=-1)
    idx = (cdf <= u.unsqueeze(-1)).sum(dim=-1)
    return idx.clamp_(max=q.shape[-1] - 1)


def sample_visible(st: RBMState, p: torch.Tensor, u: torch.Tensor,
                   ucat: Sequence[torch.Tensor]) -> torch.Tensor:
    """rbm.py:118-135.  ``u`` [B,V] uniforms for the Bernoulli part (all columns consume one),
    ``ucat[g]`` [B] the uniform of softmax group g."""
    v = (p > u).to(p.dtype)
    rows = torch.arange(p.shape[0])
    for g, (s, e) in enumerate(st.groups):
        q = p[:, s:e].clamp(1e-8, 1)
        idx = categorical_index(q, ucat[g])
        v[:, s:e] = 0.0
        v[rows, s + idx] = 1.0
    return v


def free_energy(st: RBMState, v: torch.Tensor) -> torch.Tensor:
    """F(v) = -v.b_v - sum_j softplus(b_h + vW)_j   (imdbn/utils/energy_utils.py:18-28)."""
    pre = v @ st.W + st.hb
    return -(v * st.vb).sum(dim=1) - torch.nn.functional.softplus(pre).sum(dim=1)


# --------------------------------------------------------------------------------------------
# CD-k  (rbm.py:180-227)
# --------------------------------------------------------------------------------------------
def cd_draws(k_step: int) -> Tuple[int, int, int]:
    """Draw indices of CD step ``k_step``: (U[B,V], categorical, U[B,H])."""
    return 1 + 3 * k_step, 2 + 3 * k_step, 3 + 3 * k_step


def cd_statistics(st: RBMState, data: torch.Tensor, k: int, fld: RandomField, row0: int = 0):
    """Positive/negative phase of ``train_epoch`` (rbm.py:198-209) without the update.

    Draws: 0 = U[B,H] for h0; step s: 1+3s = U[B,V], 2+3s = categorical (col = group),
    3+3s = U[B,H] (the last one is drawn by the reference but never used).
    Returns dict with the sufficient statistics as plain sums over the rows given.
    """
    B = data.shape[0]
    pos_h = hidden_probs(st, data)
    pos_assoc = data.T @ pos_h
    h = (pos_h > _t(fld.uniform(0, B, st.H, row0), data)).to(data.dtype)
    v = data
    v_prob = data
    h_prob = pos_h
    for s in range(int(k)):
        dv, dc, dh = cd_draws(s)
        v_prob = visible_probs(st, h)
        ucat = [_t(fld.cat_uniform(dc, B, g, row0), data) for g in range(len(st.groups))]
        v = sample_visible(st, v_prob, _t(fld.uniform(dv, B, st.V, row0), data), ucat)
        h_prob = hidden_probs(st, v)
        h = (h_prob > _t(fld.uniform(dh, B, st.H, row0), data)).to(data.dtype)
    neg_assoc = v.T @ h_prob
    return dict(pos_h=pos_h, h_prob=h_prob, v=v, v_prob=v_prob, h_last=h,
                dS=pos_assoc - neg_assoc,
                dh=pos_h.sum(0) - h_prob.sum(0), dv=data.sum(0) - v.sum(0),
                pos_h_sum=pos_h.sum(0), sq_err=((data - v_prob) ** 2).sum())


def apply_update(st: RBMState, dS, dh, dv, bsz: int, lr: float, mom: float,
                 pos_h_mean: Optional[torch.Tensor] = None) -> None:
    """rbm.py:211-224 (and 474-481 with lr pre-multiplied by aux_lr_mult): in place."""
    st.Wm.mul_(mom).add_(lr * (dS / bsz - st.weight_decay * st.W))
    st.W.add_(st.Wm)
    st.hbm.mul_(mom).add_(lr * dh / bsz)
    if pos_h_mean is not None:
        st.hbm.add_(-lr * (pos_h_mean - st.sparsity_factor))
    st.hb.add_(st.hbm)
    st.vbm.mul_(mom).add_(lr * dv / bsz)
    st.vb.add_(st.vbm)


def cd_train(st: RBMState, data: torch.Tensor, epoch: int, k: int, fld: RandomField):
    """``RBM.train_epoch`` (rbm.py:180-227).  Returns (loss, stats)."""
    lr, mom = lr_and_momentum(st, epoch)
    B = data.shape[0]
    s = cd_statistics(st, data, k, fld)
    # rbm.py:212 computes lr * ((pos - neg)/bsz - wd*W); rbm.py:216 lr * (dh) / bsz
    st.Wm.mul_(mom).add_(lr * (s["dS"] / B - st.weight_decay * st.W))
    st.W.add_(st.Wm)
    st.hbm.mul_(mom).add_(lr * s["dh"] / B)
    if st.sparsity:
        st.hbm.add_(-lr * (s["pos_h"].mean(0) - st.sparsity_factor))
    st.hb.add_(st.hbm)
    st.vbm.mul_(mom).add_(lr * s["dv"] / B)
    st.vb.add_(st.vbm)
    loss = torch.mean((data - s["v_prob"]) ** 2)
    return loss, s


# --------------------------------------------------------------------------------------------
# schedules (rbm.py:229-238, 338-341, 362)
# --------------------------------------------------------------------------------------------
def lin_schedule(t: int, t_max: int, start: float, end: float) -> float:
    """rbm.py:229-234."""
    if t_max <= 1:
        return float(end)
    a = min(max(t / (t_max - 1), 0.0), 1.0)
    return float(start + (end - start) * a)


def noisy_mf_schedule(n_steps: int, T0=3.0, T1=1.0, sigma0=0.9, sharpen_last=3, T_cold_plus=0.9,
                      eta0=0.15):
    """Per-step (T_t, sigma_t, eta_t) tables of ``noisy_meanfield_annealed`` (rbm.py:337-341,362)."""
    T, S, E = [], [], []
    n = int(n_steps)
    for t in range(n):
        Tt = lin_schedule(t, n, T0, T1)
        if (n - t) <= max(1, int(sharpen_last)):
            Tt = T_cold_plus
        frac = max(0.0, 1.0 - (t / max(1, n - 1)))
        T.append(max(1e-6, Tt))
        S.append(sigma0 * frac)
        E.append(eta0 * frac)
    return T, S, E


# --------------------------------------------------------------------------------------------
# conditional inference
# --------------------------------------------------------------------------------------------
def noisy_meanfield(st: RBMState, v_known, km, n_steps=72, T0=3.0, T1=1.0, sigma0=0.9,
                    hot_frac=0.7, sharpen_last=3, T_cold_plus=0.9, mu_pull=None,
                    fld: RandomField = None, draw0: int = 0, row0: int = 0):
    """``noisy_meanfield_annealed`` (rbm.py:300-367).  ``hot_frac`` has no effect there.

    Draws: draw0 = U[B,V] init; step t: draw0+1+2t = N[B,H], draw0+2+2t = N[B,V]
    (consumed only while sigma_t > 0).  ``mu_pull`` = (mu_k [B,Dz], eta0) or None (rbm.py:359-363).
    """
    B = v_known.shape[0]
    n = int(n_steps)
    eta0 = float(mu_pull[1]) if mu_pull is not None else 0.0
    Ts, Ss, Es = noisy_mf_schedule(n, T0, T1, sigma0, sharpen_last, T_cold_plus, eta0)
    v = v_known * km + (1 - km) * _t(fld.uniform(draw0, B, st.V, row0), v_known)
    for t in range(n):
        Tt, sig, eta = Ts[t], Ss[t], Es[t]
        h_logits = (v @ st.W + st.hb) / Tt
        if sig > 0:
            h_logits = h_logits + _t(fld.normal(draw0 + 1 + 2 * t, B, st.H, row0), v) * sig
        h_prob = torch.sigmoid(h_logits)
        v_logits = (h_prob @ st.W.T + st.vb) / Tt
        if sig > 0:
            v_logits = v_logits + _t(fld.normal(draw0 + 2 + 2 * t, B, st.V, row0), v) * sig
        v_prob = _group_softmax_(torch.sigmoid(v_logits), v_logits, st.groups)
        if mu_pull is not None:
            mu = mu_pull[0]
            Dz = mu.shape[1]
            v_prob[:, :Dz] = (1 - eta) * v_prob[:, :Dz] + eta * mu
        v = v_prob * (1 - km) + v_known * km
    return v


def cond_gibbs_draws(t: int) -> Tuple[int, int, int]:
    """Draw indices of conditional-Gibbs step t: (U[B,H], U[B,V], categorical)."""
    return 1 + 3 * t, 2 + 3 * t, 3 + 3 * t


def conditional_gibbs(st: RBMState, v_known, km, n_steps=30, sample_h=False, sample_v=False,
                      fld: RandomField = None, row0: int = 0):
    """``conditional_gibbs`` (rbm.py:369-400): n clamped sweeps, then ONE un-clamped sweep.

    Draws: 0 = U[B,V] init; step t: 1+3t = U[B,H] (if sample_h), 2+3t = U[B,V] and 3+3t =
    categorical (if sample_v).
    """
    B = v_known.shape[0]
    v = v_known * km + (1 - km) * _t(fld.uniform(0, B, st.V, row0), v_known)
    for t in range(int(n_steps)):
        dh, dv, dc = cond_gibbs_draws(t)
        h_prob = hidden_probs(st, v)
        h = (h_prob > _t(fld.uniform(dh, B, st.H, row0), v)).to(v.dtype) if sample_h else h_prob
        v_prob = visible_probs(st, h)
        v = v_prob * (1 - km) + v_known * km
        if sample_v:
            ucat = [_t(fld.cat_uniform(dc, B, g, row0), v) for g in range(len(st.groups))]
            v = sample_visible(st, v, _t(fld.uniform(dv, B, st.V, row0), v), ucat) * (1 - km) \
                + v_known * km
    return visible_probs(st, hidden_probs(st, v))


def gibbs_conditional_step(st: RBMState, v, v_known, km, sample_h=False, sample_v=False,
                           fld: RandomField = None, draw0: int = 0, row0: int = 0):
    """imdbn/utils/conditional_steps.py:15-34.  Draws: draw0 = U[B,H], draw0+1 = U[B,V],
    draw0+2 = categorical.  Returns (v_next, v_prob)."""
    B = v.shape[0]
    h_prob = hidden_probs(st, v)
    h = (h_prob > _t(fld.uniform(draw0, B, st.H, row0), v)).to(v.dtype) if sample_h else h_prob
    v_prob = visible_probs(st, h)
    if sample_v:
        ucat = [_t(fld.cat_uniform(draw0 + 2, B, g, row0), v) for g in range(len(st.groups))]
        v_next = sample_visible(st, v_prob, _t(fld.uniform(draw0 + 1, B, st.V, row0), v), ucat)
    else:
        v_next = v_prob
    return v_next * (1 - km) + v_known * km, v_prob


def backward_sample(st: RBMState, h: torch.Tensor, fld: RandomField, row0: int = 0) -> torch.Tensor:
    """``RBM.backward_sample`` (rbm.py:153-156).  Draws: 0 = U[B,V], 1 = categorical."""
    B = h.shape[0]
    p = visible_probs(st, h)
    ucat = [_t(fld.cat_uniform(1, B, g, row0), p) for g in range(len(st.groups))]
    return sample_visible(st, p, _t(fld.uniform(0, B, st.V, row0), p), ucat)


def gibbs_step(st: RBMState, v: torch.Tensor, sample_h: bool = True, sample_v: bool = True,
               fld: RandomField = None, row0: int = 0):
    """``RBM.gibbs_step`` (rbm.py:158-178): returns (v_next, v_prob, h, h_prob).
    Draws: 0 = U[B,H], 1 = U[B,V], 2 = categorical."""
    B = v.shape[0]
    h_prob = hidden_probs(st, v)
    h = (h_prob > _t(fld.uniform(0, B, st.H, row0), v)).to(v.dtype) if sample_h else h_prob
    v_prob = visible_probs(st, h)
    if sample_v:
        ucat = [_t(fld.cat_uniform(2, B, g, row0), v) for g in range(len(st.groups))]
        v_next = sample_visible(st, v_prob, _t(fld.uniform(1, B, st.V, row0), v), ucat)
    else:
        v_next = v_prob
    return v_next, v_prob, h, h_prob


def conditional_gibbs_annealed(st: RBMState, v_known, km, n_steps=40, T0=2.5, T1=1.0, sample_h_until=20,
                               sample_v_every=0, final_meanfield=True, fld: RandomField = None, row0: int = 0):
    """``RBM.conditional_gibbs_annealed`` (rbm.py:240-298).  Draws: 0 = U[B,V] init; step t: 1+3t = U[B,H]
    while t < min(n, sample_h_until); 2+3t = U[B,V] and 3+3t = categorical on the steps that sample v."""
    B = v_known.shape[0]
    n = int(n_steps)
    v = v_known * km + (1 - km) * _t(fld.uniform(0, B, st.V, row0), v_known)
    hot = int(max(0, min(n, sample_h_until)))
    for t in range(n):
        Tt = lin_schedule(t, n, T0, T1)
        if (n - t) <= 3:
            Tt = min(0.9, Tt)
        h_prob = hidden_probs(st, v, T=Tt)
        h = (h_prob > _t(fld.uniform(1 + 3 * t, B, st.H, row0), v)).to(v.dtype) if t < hot else h_prob
        v_prob = visible_probs(st, h, T=Tt)
        if t < hot and sample_v_every > 0 and t % sample_v_every == 0:
            ucat = [_t(fld.cat_uniform(3 + 3 * t, B, g, row0), v) for g in range(len(st.groups))]
            v_new = sample_visible(st, v_prob, _t(fld.uniform(2 + 3 * t, B, st.V, row0), v), ucat)
        else:
            v_new = v_prob
        v = v_new * (1 - km) + v_known * km
    if final_meanfield:
        v = visible_probs(st, hidden_probs(st, v, T=1.0), T=1.0) * (1 - km) + v_known * km
    return v


def finetune_last_layer(layers: Sequence[RBMState], batches: Sequence[torch.Tensor], epochs: int, lr_scale: float,
                        k: int, flds: Sequence[RandomField]) -> None:
    """``iMDBN.finetune_image_last_layer`` (imdbn.py:344-384): the lower layers only feed forward, the last layer
    trains with its learning rate scaled (restored afterwards); one random field per ``train_epoch`` call."""
    last = layers[-1]
    old = last.lr
    last.lr = max(1e-8, old * float(lr_scale))
    it = iter(flds)
    for ep in range(int(epochs)):
        for x in batches:
            v = x.reshape(x.shape[0], -1).float()
            for st in layers[:-1]:
                v = hidden_probs(st, v)
            cd_train(last, v, ep, k, next(it))
    last.lr = old


def cd_clamped_statistics(st: RBMState, v_known, km, k=1, cond_init_steps=50, sample_h=True,
                          sample_v=False, reclamp_negative=True, use_noisy_init=True,
                          fld: RandomField = None, row0: int = 0):
    """Positive/negative phase of ``train_epoch_clamped`` (rbm.py:442-472) without the update.

    Draws: the init chain uses its own numbering from 0 (noisy MF: 0..2n with n = max(10, c);
    conditional Gibbs: 0..3c); the negative phase continues at base = 1+2n (resp. 1+3c):
    step s: base+3s = U[B,H] (sample_h), base+3s+1 = U[B,V], base+3s+2 = categorical (sample_v).
    """
    B = v_known.shape[0]
    if use_noisy_init:
        n = max(10, int(cond_init_steps))
        v_plus = noisy_meanfield(st, v_known, km, n_steps=n, T0=3.0, T1=1.0, sigma0=0.9,
                                 hot_frac=0.7, sharpen_last=2, T_cold_plus=0.9,
                                 mu_pull=None, fld=fld, row0=row0)
        base = 1 + 2 * n
    else:
        v_plus = conditional_gibbs(st, v_known, km, n_steps=cond_init_steps, sample_h=sample_h,
                                   sample_v=sample_v, fld=fld, row0=row0)
        base = 1 + 3 * int(cond_init_steps)
    h_plus = hidden_probs(st, v_plus)
    pos_assoc = v_plus.T @ h_plus
    v_neg = v_plus.clone()
    for s in range(int(k)):
        h_prob = hidden_probs(st, v_neg)
        h = (h_prob > _t(fld.uniform(base + 3 * s, B, st.H, row0), v_neg)).to(v_neg.dtype) \
            if sample_h else h_prob
        v_prob = visible_probs(st, h)
        v_neg = v_prob * (1 - km) + v_known * km if reclamp_negative else v_prob
        if sample_v:
            ucat = [_t(fld.cat_uniform(base + 3 * s + 2, B, g, row0), v_neg)
                    for g in range(len(st.groups))]
            v_neg = sample_visible(st, v_neg, _t(fld.uniform(base + 3 * s + 1, B, st.V, row0),
                                                 v_neg), ucat)
    h_neg = hidden_probs(st, v_neg)
    neg_assoc = v_neg.T @ h_neg
    return dict(v_plus=v_plus, v_neg=v_neg, h_plus=h_plus, h_neg=h_neg,
                dS=pos_assoc - neg_assoc, dh=h_plus.sum(0) - h_neg.sum(0),
                dv=v_plus.sum(0) - v_neg.sum(0), sq_err=((v_plus - v_neg) ** 2).sum())


def cd_train_clamped(st: RBMState, v_known, km, epoch: int, k=1, cond_init_steps=50,
                     sample_h=True, sample_v=False, reclamp_negative=True, aux_lr_mult=0.3,
                     use_noisy_init=True, fld: RandomField = None):
    """``RBM.train_epoch_clamped`` (rbm.py:402-483).  Returns (loss, stats)."""
    lr, mom = lr_and_momentum(st, epoch)
    B = v_known.shape[0]
    s = cd_clamped_statistics(st, v_known, km, k, cond_init_steps, sample_h, sample_v,
                              reclamp_negative, use_noisy_init, fld)
    wd_term = st.weight_decay * st.W
    st.Wm.mul_(mom).add_(aux_lr_mult * lr * (s["dS"] / B - wd_term))
    st.W.add_(st.Wm)
    st.hbm.mul_(mom).add_(aux_lr_mult * lr * s["dh"] / B)
    st.hb.add_(st.hbm)
    st.vbm.mul_(mom).add_(aux_lr_mult * lr * s["dv"] / B)
    st.vb.add_(st.vbm)
    return torch.mean((s["v_plus"] - s["v_neg"]) ** 2), s


# --------------------------------------------------------------------------------------------
# iDBN (imdbn/models/idbn.py)
# --------------------------------------------------------------------------------------------
def idbn_represent(layers: Sequence[RBMState], x: torch.Tensor, upto: Optional[int] = None):
    """idbn.py:307-323."""
    v = x.reshape(x.shape[0], -1)
    L = len(layers) if upto is None else max(0, min(len(layers), int(upto)))
    for i in range(L):
        v = hidden_probs(layers[i], v)
    return v


def idbn_decode(layers: Sequence[RBMState], top: torch.Tensor):
    """idbn.py:346-359 (``backward`` = ``visible_probs``, rbm.py:137-151)."""
    cur = top
    for st in reversed(layers):
        cur = visible_probs(st, cur)
    return cur


def idbn_reconstruct(layers: Sequence[RBMState], x: torch.Tensor):
    """idbn.py:325-344."""
    return idbn_decode(layers, idbn_represent(layers, x))


def idbn_train_batch(layers: Sequence[RBMState], v: torch.Tensor, epoch: int, k: int,
                     fields: Sequence[RandomField]):
    """One minibatch of ``iDBN.train`` (idbn.py:199-204): every layer is updated, then the batch
    is pushed through the *updated* layer.  ``fields[i]`` is the random field of layer i's call."""
    losses = []
    for st, fld in zip(layers, fields):
        loss, _ = cd_train(st, v, epoch, k, fld)
        v = hidden_probs(st, v)
        losses.append(float(loss))
    return losses, v


# --------------------------------------------------------------------------------------------
# iMDBN (imdbn/models/imdbn.py)
# --------------------------------------------------------------------------------------------
def joint_bias_init(joint: RBMState, z_batches: Sequence[torch.Tensor],
                    y_batches: Sequence[torch.Tensor], Dz: int, K: int):
    """``init_joint_bias_from_data`` (imdbn.py:216-292) on already-represented batches.
    Sets joint.vb in place; returns (z_class_mean [K,Dz], z_class_count [K])."""
    sum_z = None
    n = 0
    counts = torch.zeros(K, dtype=joint.W.dtype)
    for z, y in zip(z_batches, y_batches):
        sum_z = z.sum(0) if sum_z is None else (sum_z + z.sum(0))
        n += z.shape[0]
        counts += y.to(counts.dtype).sum(0)
    mean_z = (sum_z / n).clamp(1e-4, 1 - 1e-4)
    pri = counts / max(1, counts.sum())
    pri = (pri + 1e-6) / (pri.sum() + 1e-6 * K)
    zc = torch.zeros(K, Dz, dtype=joint.W.dtype)
    cnt = torch.zeros(K, dtype=joint.W.dtype)
    for z, y in zip(z_batches, y_batches):
        yi = y.argmax(dim=1)
        for c in range(K):
            m = yi == c
            if m.any():
                zc[c] += z[m].sum(0)
                cnt[c] += m.sum()
    for c in range(K):
        if cnt[c] > 0:
            zc[c] /= cnt[c]
        else:
            zc[c] = mean_z.clone()
    joint.vb[:Dz] = torch.log(mean_z) - torch.log1p(-mean_z)
    joint.vb[Dz:Dz + K] = torch.log(pri)
    return zc, cnt


def cross_reconstruct(layers: Sequence[RBMState], joint: RBMState, z_img, y_onehot, steps: int,
                      z_class_mean=None, use_free_energy: bool = False, n_candidates: int = 5,
                      fld_i2t: RandomField = None, fld_t2i: RandomField = None,
                      fld_refine: Sequence[RandomField] = ()):
    """``iMDBN._cross_reconstruct`` (imdbn.py:386-488).

    IMG->TXT: conditional Gibbs with z clamped (field ``fld_i2t``).  TXT->IMG: noisy MF with the
    labels clamped and the mu-pull (field ``fld_t2i``), then ``n_candidates-1`` one-step cold
    refinements, each started from the previous candidate with a fresh random z
    (``fld_refine[c]``, draw 0 only), candidate picked by arg-min free energy when
    ``use_free_energy`` (the reference as shipped has no ``RBM.free_energy`` -> all zeros ->
    index 0).  Returns (img_from_txt, p_y_given_img, v_pick, best_idx).
    """
    B = z_img.shape[0]
    Dz = z_img.shape[1]
    K = y_onehot.shape[1]
    V = Dz + K
    vk = torch.zeros(B, V, dtype=z_img.dtype)
    km = torch.zeros_like(vk)
    vk[:, :Dz] = z_img
    km[:, :Dz] = 1.0
    v_i2t = conditional_gibbs(joint, vk, km, n_steps=steps, fld=fld_i2t)
    p_y = v_i2t[:, Dz:]

    vk = torch.zeros(B, V, dtype=z_img.dtype)
    km = torch.zeros_like(vk)
    vk[:, Dz:] = y_onehot
    km[:, Dz:] = 1.0
    mu = None
    if z_class_mean is not None:
        mu = (z_class_mean[y_onehot.argmax(dim=1)], 0.15)
    v_chain = noisy_meanfield(joint, vk, km, n_steps=steps, T0=3.0, T1=1.0, sigma0=0.9,
                              hot_frac=0.7, sharpen_last=3, T_cold_plus=0.9, mu_pull=mu,
                              fld=fld_t2i)
    cands = [v_chain]
    for c in range(n_candidates - 1):
        cands.append(noisy_meanfield(joint, cands[-1], km, n_steps=1, T0=0.9, T1=0.9, sigma0=0.0,
                                     hot_frac=0.0, sharpen_last=0, T_cold_plus=0.9, mu_pull=mu,
                                     fld=fld_refine[c]))
    if use_free_energy:
        F = torch.stack([free_energy(joint, c) for c in cands], dim=0)
    else:
        F = torch.zeros(len(cands), B, dtype=z_img.dtype)
    best = F.argmin(dim=0)
    v_pick = torch.stack([cands[int(best[b])][b] for b in range(B)], dim=0)
    img = idbn_decode(layers, v_pick[:, :Dz])
    return img, p_y, v_pick, best


# --------------------------------------------------------------------------------------------
# synthetic data of the benchmark (SURVEY 8d / BASELINE.md 4.3)
# --------------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------------
# IMG->TXT energy diagnostics (imdbn/utils/energy_utils.py)
# --------------------------------------------------------------------------------------------
def class_free_energies(joint: RBMState, z: torch.Tensor, K: int, Dz: int) -> torch.Tensor:
    """F_k(z) = F([z, e_k]), [B,Dz] -> [B,K]  (energy_utils.py:32-54)."""
    Wz, Wy = joint.W[:Dz], joint.W[Dz:Dz + K]
    z_bz = (z * joint.vb[:Dz].unsqueeze(0)).sum(dim=1, keepdim=True)
    pre = (z @ Wz + joint.hb.unsqueeze(0)).unsqueeze(1) + Wy.unsqueeze(0)
    return -(z_bz + joint.vb[Dz:Dz + K].unsqueeze(0)) - torch.nn.functional.softplus(pre).sum(dim=2)


def img2txt_lite_step(joint: RBMState, v: torch.Tensor, Dz: int, K: int) -> torch.Tensor:
    """energy_utils.py:61-90 in its deterministic configuration: plain sigmoids, z re-clamped, y = softmax of
    the SIGMOID outputs of the label units."""
    h = torch.sigmoid(v @ joint.W + joint.hb)
    v_next = torch.sigmoid(h @ joint.W.T + joint.vb)
    v_next[:, :Dz] = v[:, :Dz]
    v_next[:, Dz:Dz + K] = torch.softmax(v_next[:, Dz:Dz + K], dim=1)
    return v_next


def trace_single_img2txt(layers: Sequence[RBMState], joint: RBMState, x: torch.Tensor, gt: Optional[int], K: int,
                         steps: int = 30, eps_l1: float = 1e-3, stable_steps: int = 3, gap_thresh: float = 0.25):
    """energy_utils.py:96-196 for one flattened image x [1,D]."""
    z = idbn_represent(layers, x).clamp(1e-6, 1 - 1e-6)
    Dz = z.shape[1]
    Fk = class_free_energies(joint, z, K, Dz).squeeze(0)
    kstar = int(torch.argmin(Fk))
    Fmin = Fk[kstar]
    top2 = torch.topk(Fk, k=2, largest=False).values
    out = dict(margin_energy=float(top2[1] - top2[0]), kstar=kstar, gt=gt, p_top1=[], p_top2=[], p_gap=[], p_gt=[],
               deltaF_pred_traj=[])
    y = torch.full((1, K), 1.0 / K)
    v = torch.cat([z, y], dim=1)
    y_prev, pred_cur, streak, conv = y.clone(), int(y.argmax(dim=1)), 0, steps + 1
    for t in range(1, steps + 1):
        v = img2txt_lite_step(joint, v, Dz, K)
        y = v[:, Dz:Dz + K]
        vals, _ = y.topk(2, dim=1)
        p1, p2 = float(vals[0, 0]), float(vals[0, 1])
        out["p_top1"].append(p1); out["p_top2"].append(p2); out["p_gap"].append(p1 - p2)
        if gt is not None:
            out["p_gt"].append(float(y[0, gt]))
        pred_new = int(y.argmax(dim=1))
        streak = streak + 1 if pred_new == pred_cur else 1
        pred_cur = pred_new
        out["deltaF_pred_traj"].append(float(Fk[pred_cur] - Fmin))
        l1 = float((y - y_prev).abs().sum())
        if l1 < eps_l1 and streak >= stable_steps and (pred_cur == kstar or (p1 - p2) >= gap_thresh):
            conv = t
            break
        y_prev = y.clone()
    fe = torch.softmax(-Fk, dim=0)
    t2 = fe.topk(2).values
    out.update(steps_to_converge=conv, predT=pred_cur, fe_top1_final=float(fe.max()), fe_gap_final=float(t2[0] - t2[1]))
    return out


# --------------------------------------------------------------------------------------------
# iMDBN_BiModal (imdbn/models/imdbn_bimodal.py)
# --------------------------------------------------------------------------------------------
def bimodal_clamp(z: torch.Tensor, Dz1: int, Dz2: int, first: bool):
    """(v_known, known_mask) with one modality clamped (imdbn_bimodal.py:655-660, 675-678)."""
    B = z.shape[0]
    vk = torch.zeros(B, Dz1 + Dz2, dtype=z.dtype)
    km = torch.zeros_like(vk)
    if first:
        vk[:, :Dz1] = z; km[:, :Dz1] = 1.0
    else:
        vk[:, Dz1:] = z; km[:, Dz1:] = 1.0
    return vk, km


def bimodal_bias_init(joint0: RBMState, z1_batches, z2_batches, Dz1: int) -> None:
    """imdbn_bimodal.py:617-646: vis_bias of the first joint layer = logit(mean latent) per modality."""
    n = sum(z.shape[0] for z in z1_batches)
    if n == 0:
        return
    s1 = z1_batches[0].sum(0)
    for z in z1_batches[1:]:
        s1 = s1 + z.sum(0)
    s2 = z2_batches[0].sum(0)
    for z in z2_batches[1:]:
        s2 = s2 + z.sum(0)
    m1 = (s1 / n).clamp(1e-4, 1 - 1e-4)
    m2 = (s2 / n).clamp(1e-4, 1 - 1e-4)
    joint0.vb[:Dz1] = torch.log(m1) - torch.log1p(-m1)
    joint0.vb[Dz1:] = torch.log(m2) - torch.log1p(-m2)


def bimodal_cross_reconstruct(mod1: Sequence[RBMState], mod2: Sequence[RBMState], joint0: RBMState, z1, z2,
                              steps: int, seed: int, stream: int):
    """imdbn_bimodal.py:648-694; consumes two call numbers of the first joint layer.
    Returns (mod1_from_mod2, mod2_from_mod1)."""
    Dz1, Dz2 = z1.shape[1], z2.shape[1]
    vk, km = bimodal_clamp(z1, Dz1, Dz2, True)
    z2_from_1 = conditional_gibbs(joint0, vk, km, n_steps=steps, sample_h=True, sample_v=False,
                                  fld=RandomField(seed, stream))[:, Dz1:]
    vk, km = bimodal_clamp(z2, Dz1, Dz2, False)
    z1_from_2 = conditional_gibbs(joint0, vk, km, n_steps=steps, sample_h=True, sample_v=False,
                                  fld=RandomField(seed, stream + 1))[:, :Dz1]
    return idbn_decode(mod1, z1_from_2), idbn_decode(mod2, z2_from_1)


def bimodal_train_joint(mod1: Sequence[RBMState], mod2: Sequence[RBMState], joint: Sequence[RBMState], batches,
                        epochs: int, joint_cd: int, aux_steps: int, cross_steps: int, seeds: Sequence[int],
                        bias_batches: int = 10, warmup_epochs: int = 8, streams: Optional[List[int]] = None):
    """imdbn_bimodal.py:711-834 without the logging.  ``batches`` = list of (v1, v2) flattened fp32;
    ``seeds[i]`` = random-field seed of joint layer i, every layer counting its calls from ``streams[i]``
    (default 0).
    Returns the per-epoch records {mod1_mse, mod2_mse, cd_loss}."""
    z_of = lambda layers, v: idbn_represent(layers, v)
    bimodal_bias_init(joint[0], [z_of(mod1, a) for a, _ in batches[:bias_batches]],
                      [z_of(mod2, b) for _, b in batches[:bias_batches]], mod1[-1].H)
    streams = list(streams) if streams is not None else [0] * len(joint)
    Dz1, Dz2 = mod1[-1].H, mod2[-1].H
    kw = dict(k=3, cond_init_steps=aux_steps, sample_h=True, sample_v=False, aux_lr_mult=0.3, use_noisy_init=True)

    def clamped(vk, km, epoch, reclamp):
        cd_train_clamped(joint[0], vk, km, epoch, reclamp_negative=reclamp,
                         fld=RandomField(seeds[0], streams[0]), **kw)
        streams[0] += 1

    history = []
    for epoch in range(epochs):
        tot1 = tot2 = 0.0
        n = 0
        cd_losses = []
        for v1, v2 in batches:
            z1, z2 = z_of(mod1, v1), z_of(mod2, v2)
            vk1, km1 = bimodal_clamp(z1, Dz1, Dz2, True)
            vk2, km2 = bimodal_clamp(z2, Dz1, Dz2, False)
            if epoch < warmup_epochs:
                for _ in range(2):
                    clamped(vk1, km1, epoch, True)
                    clamped(vk2, km2, epoch, True)
            else:
                cur = torch.cat([z1, z2], 1)
                for li, st in enumerate(joint):
                    loss, _ = cd_train(st, cur, epoch, joint_cd, RandomField(seeds[li], streams[li]))
                    streams[li] += 1
                    if li == 0:
                        cd_losses.append(float(loss))
                    cur = hidden_probs(st, cur)
                clamped(vk1, km1, epoch, False)
                clamped(vk2, km2, epoch, False)
            r1, r2 = bimodal_cross_reconstruct(mod1, mod2, joint[0], z1, z2, cross_steps, seeds[0], streams[0])
            streams[0] += 2
            tot1 += float(((r1 - v1) ** 2).sum()); tot2 += float(((r2 - v2) ** 2).sum())
            n += v1.shape[0]
        history.append(dict(mod1_mse=tot1 / (n * v1.shape[1]), mod2_mse=tot2 / (n * v2.shape[1]),
                            cd_loss=float(np.mean(cd_losses)) if cd_losses else float("nan")))
    return history


def synthetic_images(n: int, d: int = 10000, p: float = 0.10, seed: int = 1234) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.rand(n, d, generator=g) < p).float()


def synthetic_labels(n: int, k: int = 32, seed: int = 1235) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    idx = torch.randint(0, k, (n,), generator=g)
    return torch.nn.functional.one_hot(idx, k).float()
