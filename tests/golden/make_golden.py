#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs ``/root/reference``; it does not exist on the GPU box):

    python tests/golden/make_golden.py

What it does
------------
* puts ``/root/reference`` on ``sys.path`` (with empty stub modules for the plotting packages the
  container lacks -- matplotlib, seaborn -- which the hot path never calls) and imports the
  reference's own ``imdbn.models.{rbm,idbn,imdbn}``;
* replaces the module-global ``torch`` seen by ``imdbn/models/rbm.py`` with a proxy that forwards
  everything except ``rand_like`` / ``randn_like`` / ``distributions.Categorical``; those three
  return numbers from ``oracle.philox.RandomField`` according to an explicit *draw plan*
  (the order in which the reference draws is SURVEY.md Appendix A.7);
* runs the reference methods on small seeded inputs and stores inputs + outputs as ``.npz``.

``tests/test_oracle_golden.py`` then checks the oracle restatement against these files (CPU), and
``tests/test_gpu_*.py`` check the CUDA path against them (GPU).
"""
from __future__ import annotations

import os
import sys
import types
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("IMDBN_REF", "/root/reference")

# ---- import the reference package first, before the repo root (which holds the drop-in
# ---- ``imdbn`` alias package) can shadow it
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors",
             "mpl_toolkits", "mpl_toolkits.mplot3d", "seaborn"):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib"].use = lambda *a, **k: None
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
sys.path.insert(0, REF)
os.environ.setdefault("WANDB_MODE", "disabled")
os.chdir(tempfile.mkdtemp(prefix="imdbn_golden_"))  # iDBN.__init__ creates logs-idbn/ in CWD

import numpy as np  # noqa: E402
import torch  # noqa: E402

import imdbn.models.rbm as ref_rbm_mod  # noqa: E402
import imdbn.models.idbn as ref_idbn_mod  # noqa: E402
import imdbn.models.imdbn as ref_imdbn_mod  # noqa: E402
from imdbn.utils.energy_utils import rbm_free_energy  # noqa: E402
import imdbn.utils.conditional_steps as ref_steps_mod  # noqa: E402
from imdbn.utils.conditional_steps import _gibbs_conditional_step  # noqa: E402

assert ref_rbm_mod.__file__.startswith(REF), ref_rbm_mod.__file__
sys.path.append(REPO)
from oracle.philox import RandomField  # noqa: E402
from oracle import rbm_oracle as O  # noqa: E402


# --------------------------------------------------------------------------------------------
# RNG injection
# --------------------------------------------------------------------------------------------
class DrawPlan:
    """FIFO of (kind, field, draw, group) the reference is expected to consume."""

    def __init__(self):
        self.q = []

    def push(self, kind, fld, draw, group=0, row0=0):
        self.q.append((kind, fld, draw, group, row0))

    def pop(self, kind):
        assert self.q, f"reference drew an unplanned '{kind}'"
        k, fld, draw, group, row0 = self.q.pop(0)
        assert k == kind, f"plan says '{k}' but the reference drew '{kind}'"
        self.row0 = row0
        return fld, draw, group

    def done(self):
        assert not self.q, f"{len(self.q)} planned draws were not consumed: {self.q[:3]}"


PLAN = DrawPlan()


class _Categorical:
    def __init__(self, probs=None, logits=None):
        assert probs is not None
        self.probs = probs

    def sample(self):
        fld, draw, group = PLAN.pop("cat")
        u = torch.from_numpy(fld.cat_uniform(draw, self.probs.shape[0], group, PLAN.row0))
        return O.categorical_index(self.probs, u)


class _TorchProxy:
    def __init__(self, real):
        self._real = real
        self.distributions = types.SimpleNamespace(Categorical=_Categorical)

    def __getattr__(self, name):
        return getattr(self._real, name)

    def rand_like(self, x):
        fld, draw, _ = PLAN.pop("u")
        return torch.from_numpy(fld.uniform(draw, x.shape[0], x.shape[1], PLAN.row0)).to(x.dtype)

    def randn_like(self, x):
        fld, draw, _ = PLAN.pop("n")
        return torch.from_numpy(fld.normal(draw, x.shape[0], x.shape[1], PLAN.row0)).to(x.dtype)


ref_rbm_mod.torch = _TorchProxy(torch)
ref_steps_mod.torch = ref_rbm_mod.torch   # _gibbs_conditional_step draws its own U[B,H]
RBM = ref_rbm_mod.RBM


# ---- draw plans (the order is the reference's; the draw numbers are our contract)
def plan_cd(fld, k, ngroups):
    PLAN.push("u", fld, 0)
    for s in range(k):
        dv, dc, dh = O.cd_draws(s)
        PLAN.push("u", fld, dv)
        for g in range(ngroups):
            PLAN.push("cat", fld, dc, g)
        PLAN.push("u", fld, dh)


def plan_noisy(fld, n, sigma0=0.9, draw0=0):
    PLAN.push("u", fld, draw0)
    _, S, _ = O.noisy_mf_schedule(n, sigma0=sigma0)
    for t in range(n):
        if S[t] > 0:
            PLAN.push("n", fld, draw0 + 1 + 2 * t)
            PLAN.push("n", fld, draw0 + 2 + 2 * t)


def plan_cond_gibbs(fld, n, sample_h, sample_v, ngroups):
    PLAN.push("u", fld, 0)
    for t in range(n):
        dh, dv, dc = O.cond_gibbs_draws(t)
        if sample_h:
            PLAN.push("u", fld, dh)
        if sample_v:
            PLAN.push("u", fld, dv)
            for g in range(ngroups):
                PLAN.push("cat", fld, dc, g)


def plan_clamped(fld, k, cond_init_steps, sample_h, sample_v, use_noisy_init, ngroups):
    if use_noisy_init:
        n = max(10, int(cond_init_steps))
        plan_noisy(fld, n)
        base = 1 + 2 * n
    else:
        plan_cond_gibbs(fld, cond_init_steps, sample_h, sample_v, ngroups)
        base = 1 + 3 * int(cond_init_steps)
    for s in range(k):
        if sample_h:
            PLAN.push("u", fld, base + 3 * s)
        if sample_v:
            PLAN.push("u", fld, base + 3 * s + 1)
            for g in range(ngroups):
                PLAN.push("cat", fld, base + 3 * s + 2, g)


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def make_rbm(V, H, seed, groups=None, **kw):
    hyper = dict(learning_rate=0.1, weight_decay=1e-4, momentum=0.5, dynamic_lr=True,
                 final_momentum=0.95)
    hyper.update(kw)
    r = RBM(V, H, softmax_groups=groups, **hyper).to("cpu")
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        r.W.copy_(torch.randn(V, H, generator=g) / np.sqrt(V) * 3.0)   # larger than init: makes
        r.hid_bias.copy_(torch.randn(H, generator=g) * 0.3)            # probabilities non-trivial
        r.vis_bias.copy_(torch.randn(V, generator=g) * 0.3)
    return r


def params_of(r, prefix=""):
    return {prefix + "W": r.W.detach().numpy().copy(), prefix + "hb": r.hid_bias.detach().numpy().copy(),
            prefix + "vb": r.vis_bias.detach().numpy().copy(), prefix + "Wm": r.W_m.numpy().copy(),
            prefix + "hbm": r.hb_m.numpy().copy(), prefix + "vbm": r.vb_m.numpy().copy()}


def binary(B, V, seed, p=0.3):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(B, V, generator=g) < p).float()


def unit(B, V, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, V, generator=g)


def onehot(B, K, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.one_hot(torch.randint(0, K, (B,), generator=g), K).float()


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"wrote {name}.npz  ({sum(a.nbytes for a in out.values())/1024:.1f} KiB raw)")


SEED = 20240611


# --------------------------------------------------------------------------------------------
# cases
# --------------------------------------------------------------------------------------------
def case_passes():
    """forward / visible_probs / backward(return_logits) / free energy at T=1 and T=2.5."""
    r = make_rbm(37, 19, 1, groups=[(30, 37)])
    v = unit(9, 37, 2)
    h = unit(9, 19, 3)
    save("passes", **params_of(r, "in_"), groups=np.array(r.softmax_groups), v=v, h=h,
         up_T1=r.forward(v), up_T25=r.forward(v, T=2.5),
         down_T1=r.visible_probs(h), down_T25=r.visible_probs(h, T=2.5),
         logits=r.backward(h, return_logits=True), backward=r.backward(h),
         free_energy=rbm_free_energy(r, v))


def case_train_epoch():
    """train_epoch: plain (CD-1, CD-3, two epochs straddling the momentum switch, sparsity) and
    with a softmax group (CD-2)."""
    for tag, V, H, B, groups, cds, kw in [
        ("cd_plain", 48, 24, 8, None, [1, 3, 1], {}),
        ("cd_sparse", 33, 17, 5, None, [1, 2], dict(sparsity=True, sparsity_factor=0.1)),
        ("cd_group", 26, 12, 7, [(20, 26)], [2, 1], {}),
    ]:
        r = make_rbm(V, H, 11, groups=groups, **kw)
        data = binary(B, V, 12) if groups is None else torch.cat(
            [unit(B, 20, 12), onehot(B, 6, 13)], dim=1)
        out = dict(params_of(r, "in_"), data=data, seed=SEED, cds=np.array(cds),
                   groups=np.array(groups or []).reshape(-1, 2),
                   epochs=np.array([0, 7, 3][:len(cds)]),
                   hyper=np.array([r.lr, r.weight_decay, r.momentum, r.final_momentum,
                                   float(r.dynamic_lr), float(r.sparsity), r.sparsity_factor]))
        for i, (cd, ep) in enumerate(zip(cds, out["epochs"])):
            fld = RandomField(SEED, i)
            plan_cd(fld, cd, len(groups or []))
            loss = r.train_epoch(data, int(ep), 10, CD=cd)
            PLAN.done()
            out.update(params_of(r, f"step{i}_"))
            out[f"step{i}_loss"] = loss
        save(tag, **out)


def case_noisy_mf():
    r = make_rbm(26, 12, 21, groups=[(20, 26)])
    B, Dz = 6, 20
    y = onehot(B, 6, 22)
    vk = torch.zeros(B, 26); km = torch.zeros(B, 26)
    vk[:, Dz:] = y; km[:, Dz:] = 1.0
    mu = unit(B, Dz, 23)
    out = dict(params_of(r, "in_"), groups=np.array(r.softmax_groups), v_known=vk, km=km, mu=mu,
               seed=SEED)
    # (a) default schedule, n=12, with mu-pull
    fld = RandomField(SEED, 0); plan_noisy(fld, 12)
    r._mu_pull = {"mu_k": mu, "eta0": 0.15}
    out["a"] = r.noisy_meanfield_annealed(vk, km, n_steps=12, T0=3.0, T1=1.0, sigma0=0.9,
                                          hot_frac=0.7, sharpen_last=3, T_cold_plus=0.9)
    PLAN.done()
    # (b) one cold step, no noise, mu-pull fully on (the refinement call of _cross_reconstruct)
    fld = RandomField(SEED, 1); plan_noisy(fld, 1, sigma0=0.0)
    out["b"] = r.noisy_meanfield_annealed(out["a"], km, n_steps=1, T0=0.9, T1=0.9, sigma0=0.0,
                                          hot_frac=0.0, sharpen_last=0, T_cold_plus=0.9)
    PLAN.done()
    # (c) no mu-pull, z clamped instead of y, sharpen_last=2, n=10
    r._mu_pull = None
    vk2 = torch.zeros(B, 26); km2 = torch.zeros(B, 26)
    vk2[:, :Dz] = unit(B, Dz, 24); km2[:, :Dz] = 1.0
    fld = RandomField(SEED, 2); plan_noisy(fld, 10)
    out["c"] = r.noisy_meanfield_annealed(vk2, km2, n_steps=10, sharpen_last=2)
    PLAN.done()
    out["v_known2"] = vk2; out["km2"] = km2
    save("noisy_mf", **out)


def case_cond_gibbs():
    r = make_rbm(26, 12, 31, groups=[(20, 26)])
    B, Dz = 6, 20
    vk = torch.zeros(B, 26); km = torch.zeros(B, 26)
    vk[:, :Dz] = unit(B, Dz, 32); km[:, :Dz] = 1.0
    out = dict(params_of(r, "in_"), groups=np.array(r.softmax_groups), v_known=vk, km=km, seed=SEED)
    for i, (n, sh, sv) in enumerate([(7, False, False), (5, True, False), (4, True, True),
                                     (0, False, False)]):
        fld = RandomField(SEED, i); plan_cond_gibbs(fld, n, sh, sv, 1)
        out[f"out{i}"] = r.conditional_gibbs(vk, km, n_steps=n, sample_h=sh, sample_v=sv)
        PLAN.done()
    out["cfg"] = np.array([(7, 0, 0), (5, 1, 0), (4, 1, 1), (0, 0, 0)])
    # single conditional step (utils/conditional_steps.py:15-34)
    v0 = unit(B, 26, 33)
    fld = RandomField(SEED, 9)
    PLAN.push("u", fld, 0); PLAN.push("u", fld, 1); PLAN.push("cat", fld, 2, 0)
    vn, vp = _gibbs_conditional_step(r, v0, vk, km, sample_h=True, sample_v=True)
    PLAN.done()
    vn2, vp2 = _gibbs_conditional_step(r, v0, vk, km)
    out.update(step_v0=v0, step_next=vn, step_prob=vp, step_next_mf=vn2, step_prob_mf=vp2)
    save("cond_gibbs", **out)


def case_clamped():
    B, Dz = 6, 20
    y = onehot(B, 6, 42)
    vk = torch.zeros(B, 26); km = torch.zeros(B, 26)
    vk[:, Dz:] = y; km[:, Dz:] = 1.0
    cfgs = [  # (CD, cond_init_steps, sample_h, sample_v, reclamp, noisy, epoch)
        (1, 12, False, False, True, True, 0),
        (1, 5, False, False, False, True, 9),
        (3, 4, True, False, True, True, 2),
        (2, 3, True, True, True, False, 6),
    ]
    r = make_rbm(26, 12, 41, groups=[(20, 26)])
    out = dict(params_of(r, "in_"), groups=np.array(r.softmax_groups), v_known=vk, km=km,
               seed=SEED, cfg=np.array([[int(x) for x in c] for c in cfgs]),
               hyper=np.array([r.lr, r.weight_decay, r.momentum, r.final_momentum,
                               float(r.dynamic_lr), 0.0, r.sparsity_factor]))
    for i, (cd, c, sh, sv, rc, noisy, ep) in enumerate(cfgs):
        fld = RandomField(SEED, i)
        plan_clamped(fld, cd, c, sh, sv, noisy, 1)
        loss = r.train_epoch_clamped(vk, km, ep, 10, CD=cd, cond_init_steps=c, sample_h=sh,
                                     sample_v=sv, reclamp_negative=rc, aux_lr_mult=0.3,
                                     use_noisy_init=noisy)
        PLAN.done()
        out.update(params_of(r, f"step{i}_"))
        out[f"step{i}_loss"] = loss
    save("cd_clamped", **out)


def _loader(x, y, bs):
    ds = torch.utils.data.TensorDataset(x, y)
    return torch.utils.data.DataLoader(ds, batch_size=bs, shuffle=False)


PARAMS = dict(LEARNING_RATE=0.1, WEIGHT_PENALTY=1e-4, INIT_MOMENTUM=0.5, FINAL_MOMENTUM=0.95,
              LEARNING_RATE_DYNAMIC=True, CD=1, JOINT_LEARNING_RATE=0.04, JOINT_CD=1,
              CROSS_GIBBS_STEPS=6, JOINT_AUX_COND_STEPS=4, SPARSITY=True, SPARSITY_FACTOR=0.1)


def _seed_idbn(model, base):
    for i, r in enumerate(model.layers):
        g = torch.Generator().manual_seed(base + i)
        with torch.no_grad():
            r.W.copy_(torch.randn(r.num_visible, r.num_hidden, generator=g) / np.sqrt(r.num_visible))


def case_idbn():
    """iDBN.train (2 epochs x 3 batches incl. a ragged one), represent / reconstruct / decode."""
    N, D = 20, 40
    x = binary(N, D, 51).view(N, 1, 5, 8)
    y = onehot(N, 4, 52)
    dl = _loader(x, y, 8)
    m = ref_idbn_mod.iDBN([D, 20, 10], dict(PARAMS), dl, dl, torch.device("cpu"))
    _seed_idbn(m, 60)
    out = dict(x=x, y=y, seed=SEED, batch=8, epochs=2)
    for i, r in enumerate(m.layers):
        out.update(params_of(r, f"in_l{i}_"))
    # stream convention: every RBM counts its own stochastic calls from 0
    counters = [0] * len(m.layers)
    for ep in range(2):
        for b in range(3):
            for li, r in enumerate(m.layers):
                plan_cd(RandomField(SEED + li, counters[li]), 1, 0)
                counters[li] += 1
    m.train(2)
    PLAN.done()
    for i, r in enumerate(m.layers):
        out.update(params_of(r, f"out_l{i}_"))
    out["represent"] = m.represent(x)
    out["represent1"] = m.represent(x, upto_layer=1)
    out["reconstruct"] = m.reconstruct(x)
    out["decode"] = m.decode(unit(5, 10, 53))
    out["decode_in"] = unit(5, 10, 53)
    save("idbn", **out)


def case_imdbn():
    """init_joint_bias_from_data, _cross_reconstruct (with and without a free-energy hook),
    represent, and a short train_joint (warm-up epochs + main-phase epochs)."""
    N, D, K = 20, 40, 4
    x = binary(N, D, 71).view(N, 1, 5, 8)
    y = onehot(N, K, 72)
    dl = _loader(x, y, 8)
    dev = torch.device("cpu")
    m = ref_imdbn_mod.iMDBN([D, 20, 10], 8, params=dict(PARAMS), dataloader=dl, val_loader=dl,
                            device=dev, num_labels=K)
    _seed_idbn(m.image_idbn, 80)
    jr = m.joint_rbm
    g = torch.Generator().manual_seed(90)
    with torch.no_grad():
        jr.W.copy_(torch.randn(jr.num_visible, jr.num_hidden, generator=g) * 0.5)
        jr.hid_bias.copy_(torch.randn(jr.num_hidden, generator=g) * 0.2)
    out = dict(x=x, y=y, seed=SEED, batch=8, K=K)
    for i, r in enumerate(m.image_idbn.layers):
        out.update(params_of(r, f"in_l{i}_"))
    out.update(params_of(jr, "in_joint_"))

    m.init_joint_bias_from_data(n_batches=2)
    out["bias_vb"] = jr.vis_bias.detach().clone()
    out["z_class_mean"] = m.z_class_mean.clone()
    out["z_class_count"] = m.z_class_count.clone()
    out["represent"] = m.represent((x, y))

    z = m.image_idbn.represent(x)
    JSEED = SEED + 100
    stream = 0

    def plan_cross(stream, steps):
        plan_cond_gibbs(RandomField(JSEED, stream), steps, False, False, 1)
        plan_noisy(RandomField(JSEED, stream + 1), steps)
        for c in range(4):
            plan_noisy(RandomField(JSEED, stream + 2 + c), 1, sigma0=0.0)
        return stream + 6

    stream = plan_cross(stream, 6)
    img, py = m._cross_reconstruct(z, y, steps=6)
    PLAN.done()
    out["cross_img"] = img; out["cross_py"] = py
    # with the free-energy hook installed: best-of-K becomes live
    RBM.free_energy = rbm_free_energy
    stream = plan_cross(stream, 6)
    img, py = m._cross_reconstruct(z, y, steps=6)
    PLAN.done()
    del RBM.free_energy
    out["cross_img_fe"] = img; out["cross_py_fe"] = py

    # ---- train_joint: 10 epochs (8 warm-up + 2 main) x 3 batches; resets the stream numbering
    # of the joint RBM to what the drop-in does: continue counting.
    epochs, nb = 10, 3
    aux = PARAMS["JOINT_AUX_COND_STEPS"]
    # init_joint_bias_from_data draws nothing
    for ep in range(epochs):
        for b in range(nb):
            if ep < 8:
                for _ in range(2):
                    plan_clamped(RandomField(JSEED, stream), 1, aux, False, False, True, 1); stream += 1
            else:
                plan_cd(RandomField(JSEED, stream), PARAMS["JOINT_CD"], 1); stream += 1
                plan_clamped(RandomField(JSEED, stream), 1, aux, False, False, True, 1); stream += 1
                if b % 50 == 0:
                    plan_clamped(RandomField(JSEED, stream), 1, aux, False, False, True, 1); stream += 1
            stream = plan_cross(stream, PARAMS["CROSS_GIBBS_STEPS"])
    m.train_joint(epochs)
    PLAN.done()
    out.update(params_of(jr, "out_joint_"))
    out["train_epochs"] = epochs
    out["final_stream"] = stream
    save("imdbn", **out)

    # ---- convergence traces of utils/conditional_steps.py on the trained model (one sample at a time,
    # as the reference does); IMG->TXT draws one U[1,V] for the init, TXT->IMG draws nothing.
    from imdbn.utils.conditional_steps import trace_img2txt_cross, trace_txt2img_cross
    tr = dict(seed=JSEED, stream0=stream, n=4, max_steps=12)
    for i in range(4):
        PLAN.push("u", RandomField(JSEED, stream + i), 0)
        a = trace_img2txt_cross(m, x[i:i + 1], y[i:i + 1], max_steps=12, eps_l1=1e-2, stable_steps=2,
                                gap_thresh=0.05)
        PLAN.done()
        b = trace_txt2img_cross(m, x[i:i + 1], y[i:i + 1], max_steps=12, eps_z=5e-2, mse_tol=1e-4, patience=2)
        tr[f"a{i}_steps"] = a["steps_to_converge"]; tr[f"a{i}_p_top1"] = np.array(a["p_top1"])
        tr[f"a{i}_p_gap"] = np.array(a["p_gap"]); tr[f"a{i}_l1"] = np.array(a["l1"])
        tr[f"a{i}_top1_idx"] = np.array(a["top1_idx"]); tr[f"a{i}_predT"] = a["predT"]
        tr[f"a{i}_p_gt"] = np.array(a["p_gt"]); tr[f"a{i}_gt_idx"] = a["gt_idx"]
        tr[f"b{i}_steps"] = b["steps_to_converge"]; tr[f"b{i}_z_l2"] = np.array(b["z_l2"])
        tr[f"b{i}_image_mse"] = np.array(b["image_mse"]); tr[f"b{i}_best_mse"] = b["best_mse"]
    for i, r in enumerate(m.image_idbn.layers):
        tr.update(params_of(r, f"l{i}_"))
    tr.update(params_of(jr, "joint_"))
    tr["z_class_mean"] = m.z_class_mean.clone(); tr["x"] = x; tr["y"] = y
    save("traces", **tr)

    # ---- a checkpoint written by the reference itself (pickle of reference objects, CPU tensors):
    # the drop-in must load it through the `imdbn` alias package
    m.image_idbn.save_model(os.path.join(HERE, "ref_idbn.pkl"))


def case_energy():
    """utils/energy_utils.py: class_free_energies for a batch and trace_single_img2txt for four cases on a
    seeded (untrained) iMDBN whose joint RBM has H = 32 hidden units."""
    from imdbn.utils.energy_utils import class_free_energies, trace_single_img2txt
    N, D, K = 12, 40, 4
    x = binary(N, D, 151).view(N, 1, 5, 8)
    y = onehot(N, K, 152)
    dl = _loader(x, y, 6)
    m = ref_imdbn_mod.iMDBN([D, 20, 12], 32, params=dict(PARAMS), dataloader=dl, val_loader=dl,
                            device=torch.device("cpu"), num_labels=K)
    _seed_idbn(m.image_idbn, 160)
    jr = m.joint_rbm
    g = torch.Generator().manual_seed(170)
    with torch.no_grad():
        jr.W.copy_(torch.randn(jr.num_visible, jr.num_hidden, generator=g) * 0.8)
        jr.hid_bias.copy_(torch.randn(jr.num_hidden, generator=g) * 0.3)
        jr.vis_bias.copy_(torch.randn(jr.num_visible, generator=g) * 0.3)
    out = dict(x=x, y=y, K=K)
    for i, r in enumerate(m.image_idbn.layers):
        out.update(params_of(r, f"l{i}_"))
    out.update(params_of(jr, "joint_"))
    z = m.image_idbn.represent(x)
    out["Fk"] = class_free_energies(jr, z, K, 12)
    for i in range(4):
        tr = trace_single_img2txt(m, x[i:i + 1], y[i:i + 1], steps=12, eps_l1=1e-4, stable_steps=3, gap_thresh=0.9)
        for key in ("p_top1", "p_top2", "p_gap", "p_gt", "deltaF_pred_traj"):
            out[f"t{i}_{key}"] = np.array(tr[key])
        for key in ("steps_to_converge", "kstar", "predT", "margin_energy", "fe_top1_final", "fe_gap_final", "gt"):
            out[f"t{i}_{key}"] = tr[key]
    save("energy", **out)


def case_bimodal():
    """iMDBN_BiModal (imdbn_bimodal.py): bias init, both cross directions with stochastic hidden units,
    represent through a two-layer joint stack, and train_joint (8 warm-up + 2 main epochs x 2 batches)."""
    import imdbn.models.imdbn_bimodal as ref_bi
    assert ref_bi.RBM is RBM                      # the patched reference RBM
    N, D1, D2 = 12, 36, 28
    x1 = binary(N, D1, 111).view(N, 1, 6, 6)
    x2 = binary(N, D2, 112, p=0.4).view(N, 1, 4, 7)
    dl = _loader(x1, x2, 6)
    params = dict(PARAMS, JOINT_CD=2, CROSS_GIBBS_STEPS=5, JOINT_AUX_COND_STEPS=4)
    m = ref_bi.iMDBN_BiModal([D1, 20, 12], [D2, 16, 10], [14, 8], params=params, dataloader=dl, val_loader=dl,
                             device=torch.device("cpu"))
    _seed_idbn(m.mod1_dbn, 120)
    _seed_idbn(m.mod2_dbn, 130)
    for i, r in enumerate(m.joint_layers):
        g = torch.Generator().manual_seed(140 + i)
        with torch.no_grad():
            r.W.copy_(torch.randn(r.num_visible, r.num_hidden, generator=g) * 0.5)
            r.hid_bias.copy_(torch.randn(r.num_hidden, generator=g) * 0.2)
    out = dict(x1=x1, x2=x2, seed=SEED, batch=6)
    for name, dbn in (("m1", m.mod1_dbn), ("m2", m.mod2_dbn)):
        for i, r in enumerate(dbn.layers):
            out.update(params_of(r, f"in_{name}_l{i}_"))
    for i, r in enumerate(m.joint_layers):
        out.update(params_of(r, f"in_j{i}_"))
    BSEED = SEED + 200                            # joint layer i uses seed BSEED + i, calls counted per layer
    streams = [0, 0]

    m.init_joint_bias_from_data(n_batches=2)
    out["bias_vb"] = m.joint_layers[0].vis_bias.detach().clone()
    out["represent"] = m.represent((x1, x2))

    def plan_cross(steps):
        for _ in range(2):
            plan_cond_gibbs(RandomField(BSEED, streams[0]), steps, True, False, 0)
            streams[0] += 1

    z1 = m.mod1_dbn.represent(x1.view(N, -1)); z2 = m.mod2_dbn.represent(x2.view(N, -1))
    plan_cross(5)
    r1, r2 = m._cross_reconstruct(z1, z2, steps=5)
    PLAN.done()
    out["cross_mod1"] = r1; out["cross_mod2"] = r2

    epochs, nb, aux = 10, 2, params["JOINT_AUX_COND_STEPS"]
    for ep in range(epochs):
        for b in range(nb):
            if ep < 8:
                for _ in range(4):
                    plan_clamped(RandomField(BSEED, streams[0]), 3, aux, True, False, True, 0); streams[0] += 1
            else:
                for li in range(2):
                    plan_cd(RandomField(BSEED + li, streams[li]), params["JOINT_CD"], 0); streams[li] += 1
                for _ in range(2):
                    plan_clamped(RandomField(BSEED, streams[0]), 3, aux, True, False, True, 0); streams[0] += 1
            plan_cross(params["CROSS_GIBBS_STEPS"])
    logged = []
    m.wandb_run = types.SimpleNamespace(log=lambda d: logged.append(dict(d)))
    m.val_loader = None                           # keeps the W&B plotting branches off
    m.validation_mod1 = None                      # ... and the snapshot renderer (it would draw two more chains)
    m.train_joint(epochs)
    PLAN.done()
    for i, r in enumerate(m.joint_layers):
        out.update(params_of(r, f"out_j{i}_"))
    out["train_epochs"] = epochs
    out["final_streams"] = np.array(streams)
    out["mod1_mse"] = np.array([d["cross_modality/mod1_mse"] for d in logged if "cross_modality/mod1_mse" in d])
    out["mod2_mse"] = np.array([d["cross_modality/mod2_mse"] for d in logged if "cross_modality/mod2_mse" in d])
    out["cd_loss"] = np.array([d["joint/cd_loss"] for d in logged if "joint/cd_loss" in d])
    save("bimodal", **out)


def case_rbm_extra():
    """backward_sample (rbm.py:153-156), gibbs_step (:158-178), conditional_gibbs_annealed (:240-298)."""
    r = make_rbm(26, 12, 201, groups=[(20, 26)])
    B, Dz = 6, 20
    h = unit(B, 12, 202)
    v0 = torch.cat([unit(B, Dz, 203), onehot(B, 6, 204)], dim=1)
    out = dict(params_of(r, "in_"), groups=np.array(r.softmax_groups), h=h, v0=v0, seed=SEED)
    # backward_sample: U[B,V] (draw 0), categorical (draw 1)
    fld = RandomField(SEED, 0)
    PLAN.push("u", fld, 0); PLAN.push("cat", fld, 1, 0)
    out["backward_sample"] = r.backward_sample(h)
    PLAN.done()
    # gibbs_step: U[B,H] (0), U[B,V] (1), categorical (2)
    for i, (sh, sv) in enumerate([(True, True), (True, False), (False, True), (False, False)]):
        fld = RandomField(SEED, 1 + i)
        if sh:
            PLAN.push("u", fld, 0)
        if sv:
            PLAN.push("u", fld, 1); PLAN.push("cat", fld, 2, 0)
        vn, vp, hh, hp = r.gibbs_step(v0, sample_h=sh, sample_v=sv)
        PLAN.done()
        out[f"gs{i}_v_next"] = vn; out[f"gs{i}_v_prob"] = vp; out[f"gs{i}_h"] = hh; out[f"gs{i}_h_prob"] = hp
    out["gs_cfg"] = np.array([(1, 1), (1, 0), (0, 1), (0, 0)])
    # conditional_gibbs_annealed: U[B,V] init (0); step t: U[B,H] (1+3t) while t < hot; U[B,V] (2+3t) and
    # categorical (3+3t) when v is sampled
    vk = torch.zeros(B, 26); km = torch.zeros(B, 26)
    vk[:, :Dz] = unit(B, Dz, 205); km[:, :Dz] = 1.0
    cfgs = [(8, 2.5, 1.0, 4, 2, True), (6, 2.0, 1.0, 6, 0, False), (5, 2.5, 1.0, 0, 0, True), (4, 3.0, 0.8, 2, 1, True)]
    for i, (n, T0, T1, until, every, final) in enumerate(cfgs):
        fld = RandomField(SEED, 10 + i)
        PLAN.push("u", fld, 0)
        hot = max(0, min(n, until))
        for t in range(n):
            if t < hot:
                PLAN.push("u", fld, 1 + 3 * t)
                if every > 0 and t % every == 0:
                    PLAN.push("u", fld, 2 + 3 * t); PLAN.push("cat", fld, 3 + 3 * t, 0)
        out[f"cga{i}"] = r.conditional_gibbs_annealed(vk, km, n_steps=n, T0=T0, T1=T1, sample_h_until=until,
                                                       sample_v_every=every, final_meanfield=final)
        PLAN.done()
    out["cga_cfg"] = np.array([[c[0], c[1], c[2], c[3], c[4], float(c[5])] for c in cfgs])
    out["v_known"] = vk; out["km"] = km
    save("rbm_extra", **out)


class _Recorder:
    """Stand-in for a wandb run: keeps what the reference logs (scalars / dicts only)."""

    def __init__(self):
        self.logs = []

    def log(self, d, *a, **k):
        self.logs.append(d)


def _small_imdbn(seed0, N=24, D=40, K=4, wandb_run=None):
    x = binary(N, D, seed0 + 1).view(N, 1, 5, 8)
    y = onehot(N, K, seed0 + 2)
    dl = _loader(x, y, 8)
    m = ref_imdbn_mod.iMDBN([D, 20, 10], 8, params=dict(PARAMS), dataloader=dl, val_loader=dl,
                            device=torch.device("cpu"), num_labels=K, wandb_run=wandb_run)
    _seed_idbn(m.image_idbn, seed0 + 10)
    jr = m.joint_rbm
    g = torch.Generator().manual_seed(seed0 + 20)
    with torch.no_grad():
        jr.W.copy_(torch.randn(jr.num_visible, jr.num_hidden, generator=g) * 0.8)
        jr.hid_bias.copy_(torch.randn(jr.num_hidden, generator=g) * 0.2)
    return m, x, y


def case_finetune():
    """iMDBN.finetune_image_last_layer (imdbn.py:344-384): 2 epochs x 3 batches, lr scaled by 0.3, CD-2."""
    m, x, y = _small_imdbn(300)
    out = dict(x=x, y=y, seed=SEED, batch=8)
    for i, r in enumerate(m.image_idbn.layers):
        out.update(params_of(r, f"in_l{i}_"))
    last = m.image_idbn.layers[-1]
    lr0 = float(last.lr)
    stream = 0
    for ep in range(2):
        for b in range(3):
            plan_cd(RandomField(SEED + 1, stream), 2, 0); stream += 1
    m.finetune_image_last_layer(epochs=2, lr_scale=0.3, cd_k=2)
    PLAN.done()
    assert float(last.lr) == lr0
    for i, r in enumerate(m.image_idbn.layers):
        out.update(params_of(r, f"out_l{i}_"))
    out["lr_after"] = float(last.lr)
    save("finetune", **out)


def case_panel():
    """run_and_log_cross_fixed_case (conditional_steps.py:364-387), run_and_log_cross_panel (:474-555) and
    run_and_log_z_mismatch_check (:557-646) with the plotting stubbed and a recording wandb run.
    Draw convention of the batched drop-in: ONE stream per call of the joint RBM, sample i = row i."""
    rec = _Recorder()
    # plotting is out of scope: stub the three matplotlib / wandb.Image touch points of the module
    ref_steps_mod.log_cross_case = lambda *a, **k: None
    ref_steps_mod._plot_steps_hist_with_nc = lambda *a, **k: None
    ref_steps_mod.wandb = types.SimpleNamespace(Image=lambda fig: None)
    ref_steps_mod.plt = types.SimpleNamespace(close=lambda fig: None)
    m, x, y = _small_imdbn(400, wandb_run=rec)
    jr = m.joint_rbm
    m.init_joint_bias_from_data(n_batches=3)
    out = dict(x=x, y=y, K=4)
    for i, r in enumerate(m.image_idbn.layers):
        out.update(params_of(r, f"l{i}_"))
    out.update(params_of(jr, "joint_"))
    out["z_class_mean"] = m.z_class_mean.clone()
    JSEED = SEED + 400
    kw = dict(max_steps=10)
    # ---- fixed case: one sample (the first of class 2 in the first batch that has one)
    PLAN.push("u", RandomField(JSEED, 0), 0)
    a, b = ref_steps_mod.run_and_log_cross_fixed_case(m, epoch=0, target_label=2, **kw)
    PLAN.done()
    out["fixed_img"] = m._fixed_val_case[0]; out["fixed_lbl"] = m._fixed_val_case[1]
    out["fixed_a_steps"] = a["steps_to_converge"]; out["fixed_a_p_top1"] = np.array(a["p_top1"])
    out["fixed_a_l1"] = np.array(a["l1"]); out["fixed_a_top1_idx"] = np.array(a["top1_idx"])
    out["fixed_b_steps"] = b["steps_to_converge"]; out["fixed_b_z_l2"] = np.array(b["z_l2"])
    out["fixed_b_image_mse"] = np.array(b["image_mse"])
    # ---- panel: per_class = 2
    imgs, lbls = ref_steps_mod.build_or_get_fixed_val_panel(m, per_class=2)
    n = imgs.size(0)
    for i in range(n):
        PLAN.push("u", RandomField(JSEED, 1), 0, row0=i)
    p = ref_steps_mod.run_and_log_cross_panel(m, epoch=0, per_class=2, **kw)
    PLAN.done()
    out["panel_imgs"] = imgs; out["panel_lbls"] = lbls
    out["panel_i2t_steps"] = np.array(p["img2txt"]["steps"]); out["panel_t2i_steps"] = np.array(p["txt2img"]["steps"])
    st = p["img2txt"]["stats"]
    out["panel_i2t_stats"] = np.array([st["n_total"], st["n_converged"], st["frac_converged"],
                                       -1.0 if st["mean"] is None else st["mean"],
                                       -1.0 if st["p50"] is None else st["p50"],
                                       -1.0 if st["p95"] is None else st["p95"]], dtype=np.float64)
    st = p["txt2img"]["stats"]
    out["panel_t2i_stats"] = np.array([st["n_total"], st["n_converged"], st["frac_converged"],
                                       -1.0 if st["mean"] is None else st["mean"],
                                       -1.0 if st["p50"] is None else st["p50"],
                                       -1.0 if st["p95"] is None else st["p95"]], dtype=np.float64)
    out["panel_p1_mean"] = p["img2txt"]["p1_mean"]; out["panel_gap_mean"] = p["img2txt"]["gap_mean"]
    out["panel_best_mse_mean"] = p["txt2img"]["best_mse_mean"]
    summ = [d for d in rec.logs if any(k.endswith("/summary") for k in d)]
    assert len(summ) == 1
    # ---- z mismatch check on the first validation batch (8 samples), 6 steps
    rec.logs.clear()
    for i in range(8):
        PLAN.push("u", RandomField(JSEED, 2), 0, row0=i)
    ref_steps_mod.run_and_log_z_mismatch_check(m, epoch=0, max_steps=6)
    PLAN.done()
    zl = {k.split("/")[-1]: v for d in rec.logs for k, v in d.items() if k.startswith("zcheck/")}
    out["z_img_stats"] = np.array([zl["z_img_stats"][k] for k in ("mean", "std", "q10", "q90")])
    out["z_y_stats"] = np.array([zl["z_y_stats"][k] for k in ("mean", "std", "q10", "q90")])
    out["z_cosine_mean"] = zl["cosine_mean"]
    out["seed_joint"] = JSEED
    save("panel", **out)


def case_probe():
    """probe_utils.py: quantile binning (:151-166), bin names (:169-177), stratified split (:170-189) and the linear probe
    with early stopping (:195-263) run by the reference on seeded inputs; the probe's ``nn.Linear`` is drawn from the
    torch generator seeded here (the drop-in creates its layer on the host in the same way)."""
    import imdbn.utils.probe_utils as ref_probe
    assert ref_probe.__file__.startswith(REF)
    g = torch.Generator().manual_seed(77)
    out = {}
    # continuous feature, a discrete one with ties (forces the jitter on equal edges), and class labels
    vals = {"cont": torch.rand(240, generator=g) * 7.0,
            "ties": torch.randint(0, 3, (240,), generator=g).float(),
            "labels": torch.randint(0, 8, (240,), generator=g).float()}
    for key, v in vals.items():
        y, edges = ref_probe.make_bin_labels(v.clone(), n_bins=5)
        tr, te = ref_probe.stratified_split(y, test_size=0.2, rng_seed=42)
        out[key + "_values"] = v.numpy()
        out[key + "_bins"] = y.numpy()
        out[key + "_edges"] = edges.numpy()
        out[key + "_train_idx"] = np.asarray(tr, dtype=np.int64)
        out[key + "_test_idx"] = np.asarray(te, dtype=np.int64)
        out[key + "_names"] = np.asarray(ref_probe._format_bin_names(edges, precision=4))
    # linear probe on embeddings whose first coordinates carry the class
    N, D, C = 300, 24, 5
    y = torch.randint(0, C, (N,), generator=g)
    X = torch.randn(N, D, generator=g) * 0.7
    X[torch.arange(N), y] += 2.0
    tr, te = ref_probe.stratified_split(y, test_size=0.2, rng_seed=42)
    torch.manual_seed(1234)
    acc, y_true, y_pred = ref_probe.train_linear_classifier(X[tr].numpy(), y[tr].numpy(), X[te].numpy(), y[te].numpy(),
                                                            device=torch.device("cpu"), n_classes=C, max_steps=300,
                                                            lr=1e-2, weight_decay=0.0, patience=20, min_delta=0.0)
    out.update(probe_X=X.numpy(), probe_y=y.numpy(), probe_train_idx=np.asarray(tr), probe_test_idx=np.asarray(te),
               probe_acc=np.float64(acc), probe_y_true=np.asarray(y_true), probe_y_pred=np.asarray(y_pred),
               probe_seed=np.int64(1234))
    np.savez_compressed(os.path.join(HERE, "probe.npz"), **out)
    print("probe.npz", len(out), "arrays, probe acc", acc)


if __name__ == "__main__":
    torch.set_num_threads(1)  # fixtures should not depend on the thread count of this machine
    case_passes()
    case_train_epoch()
    case_noisy_mf()
    case_cond_gibbs()
    case_clamped()
    case_idbn()
    case_imdbn()
    case_bimodal()
    case_energy()
    case_rbm_extra()
    case_finetune()
    case_panel()
    case_probe()
