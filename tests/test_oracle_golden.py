"""Pins the oracle: the restatement in ``oracle/rbm_oracle.py`` must reproduce the outputs that the
UNMODIFIED reference produced for the same inputs and the same injected random numbers
(fixtures written by ``tests/golden/make_golden.py``).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import rbm_oracle as O
from oracle.philox import KAT, RandomField, philox4x32_10

# The oracle repeats the reference's torch-CPU operations in the same order, so the match is
# expected to be exact on the machine that wrote the fixtures; the tolerance only absorbs a
# different BLAS blocking on another host.
TOL = dict(rtol=2e-6, atol=2e-7)


def T(a):
    return torch.from_numpy(np.array(a))


def state_from(g, prefix, groups=(), hyper=None):
    kw = {}
    if hyper is not None:
        lr, wd, mom, fmom, dyn, sp, spf = [float(x) for x in hyper]
        kw = dict(lr=lr, weight_decay=wd, momentum=mom, final_momentum=fmom, dynamic_lr=bool(dyn),
                  sparsity=bool(sp), sparsity_factor=spf)
    return O.RBMState(T(g[prefix + "W"]), T(g[prefix + "hb"]), T(g[prefix + "vb"]),
                      T(g[prefix + "Wm"]), T(g[prefix + "hbm"]), T(g[prefix + "vbm"]),
                      groups=[tuple(int(x) for x in r) for r in np.array(groups).reshape(-1, 2)],
                      **kw)


def check_params(st, g, prefix):
    for n in ("W", "hb", "vb", "Wm", "hbm", "vbm"):
        torch.testing.assert_close(getattr(st, n), T(g[prefix + n]), msg=lambda m: f"{prefix}{n}: {m}", **TOL)


def test_philox_known_answers():
    for ctr, key, out in KAT:
        got = philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == out


def test_uniform_range_and_row_offset():
    f = RandomField(7, 3)
    u = f.uniform(2, 64, 33)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
    # a shard starting at global row 16 sees the same numbers as rows 16.. of the full draw
    np.testing.assert_array_equal(f.uniform(2, 8, 33, row0=16), u[16:24])
    n = f.normal(5, 512, 512)
    assert abs(float(n.mean())) < 0.01 and abs(float(n.std()) - 1.0) < 0.01


def test_passes():
    g = load_golden("passes")
    st = state_from(g, "in_", g["groups"])
    v, h = T(g["v"]), T(g["h"])
    torch.testing.assert_close(O.hidden_probs(st, v), T(g["up_T1"]), **TOL)
    torch.testing.assert_close(O.hidden_probs(st, v, 2.5), T(g["up_T25"]), **TOL)
    torch.testing.assert_close(O.visible_probs(st, h), T(g["down_T1"]), **TOL)
    torch.testing.assert_close(O.visible_probs(st, h, 2.5), T(g["down_T25"]), **TOL)
    torch.testing.assert_close(O.visible_logits(st, h), T(g["logits"]), **TOL)
    torch.testing.assert_close(O.free_energy(st, v), T(g["free_energy"]), **TOL)
    s, e = st.groups[0]
    assert torch.allclose(O.visible_probs(st, h)[:, s:e].sum(1), torch.ones(h.shape[0]), atol=1e-6)


@pytest.mark.parametrize("name", ["cd_plain", "cd_sparse", "cd_group"])
def test_train_epoch(name):
    g = load_golden(name)
    st = state_from(g, "in_", g["groups"], g["hyper"])
    data = T(g["data"])
    for i, (cd, ep) in enumerate(zip(g["cds"], g["epochs"])):
        loss, _ = O.cd_train(st, data, int(ep), int(cd), RandomField(int(g["seed"]), i))
        check_params(st, g, f"step{i}_")
        torch.testing.assert_close(loss, T(g[f"step{i}_loss"]), **TOL)


def test_noisy_meanfield():
    g = load_golden("noisy_mf")
    st = state_from(g, "in_", g["groups"])
    seed = int(g["seed"])
    vk, km, mu = T(g["v_known"]), T(g["km"]), T(g["mu"])
    a = O.noisy_meanfield(st, vk, km, n_steps=12, mu_pull=(mu, 0.15), fld=RandomField(seed, 0))
    torch.testing.assert_close(a, T(g["a"]), **TOL)
    b = O.noisy_meanfield(st, T(g["a"]), km, n_steps=1, T0=0.9, T1=0.9, sigma0=0.0, hot_frac=0.0,
                          sharpen_last=0, T_cold_plus=0.9, mu_pull=(mu, 0.15),
                          fld=RandomField(seed, 1))
    torch.testing.assert_close(b, T(g["b"]), **TOL)
    c = O.noisy_meanfield(st, T(g["v_known2"]), T(g["km2"]), n_steps=10, sharpen_last=2,
                          fld=RandomField(seed, 2))
    torch.testing.assert_close(c, T(g["c"]), **TOL)


def test_schedule_known_answers():
    # SURVEY A.5
    Ts, Ss, Es = O.noisy_mf_schedule(50, sharpen_last=3)
    assert Ts[0] == 3.0 and abs(Ts[1] - 2.9591836) < 1e-6 and abs(Ts[46] - 1.1224489) < 1e-6
    assert Ts[47] == Ts[48] == Ts[49] == 0.9
    assert Ss[0] == 0.9 and abs(Ss[48] - 0.9 / 49) < 1e-9 and Ss[49] == 0.0
    Ts, Ss, _ = O.noisy_mf_schedule(30, sharpen_last=2)
    assert Ts[28] == Ts[29] == 0.9 and Ts[27] != 0.9 and Ss[29] == 0.0
    Ts, Ss, Es = O.noisy_mf_schedule(1, T0=0.9, T1=0.9, sigma0=0.0, sharpen_last=0)
    assert Ts == [0.9] and Ss == [0.0] and Es == [0.15]


def test_conditional_gibbs():
    g = load_golden("cond_gibbs")
    st = state_from(g, "in_", g["groups"])
    seed = int(g["seed"])
    vk, km = T(g["v_known"]), T(g["km"])
    for i, (n, sh, sv) in enumerate(g["cfg"]):
        out = O.conditional_gibbs(st, vk, km, int(n), bool(sh), bool(sv), fld=RandomField(seed, i))
        torch.testing.assert_close(out, T(g[f"out{i}"]), **TOL)
    # n_steps = 0 is one un-clamped sweep from the random init (SURVEY 8c known answer)
    f = RandomField(seed, 3)
    v0 = vk * km + (1 - km) * torch.from_numpy(f.uniform(0, vk.shape[0], vk.shape[1]))
    torch.testing.assert_close(O.visible_probs(st, O.hidden_probs(st, v0)), T(g["out3"]), **TOL)
    vn, vp = O.gibbs_conditional_step(st, T(g["step_v0"]), vk, km, True, True,
                                      fld=RandomField(seed, 9))
    torch.testing.assert_close(vn, T(g["step_next"]), **TOL)
    torch.testing.assert_close(vp, T(g["step_prob"]), **TOL)
    vn, vp = O.gibbs_conditional_step(st, T(g["step_v0"]), vk, km)
    torch.testing.assert_close(vn, T(g["step_next_mf"]), **TOL)


def test_train_epoch_clamped():
    g = load_golden("cd_clamped")
    st = state_from(g, "in_", g["groups"], g["hyper"])
    vk, km = T(g["v_known"]), T(g["km"])
    for i, (cd, c, sh, sv, rc, noisy, ep) in enumerate(g["cfg"]):
        loss, _ = O.cd_train_clamped(st, vk, km, int(ep), int(cd), int(c), bool(sh), bool(sv),
                                     bool(rc), 0.3, bool(noisy),
                                     fld=RandomField(int(g["seed"]), i))
        check_params(st, g, f"step{i}_")
        torch.testing.assert_close(loss, T(g[f"step{i}_loss"]), **TOL)


IDBN_HYPER = dict(lr=0.1, weight_decay=1e-4, momentum=0.5, final_momentum=0.95, dynamic_lr=True)


def idbn_layers(g, prefix, n=2, sparsity_last=True):
    layers = []
    for i in range(n):
        st = state_from(g, f"{prefix}l{i}_")
        for k, v in IDBN_HYPER.items():
            setattr(st, k, v)
        st.sparsity = sparsity_last and i == n - 1       # idbn.py:158
        st.sparsity_factor = 0.1
        layers.append(st)
    return layers


def test_idbn_train_and_chains():
    g = load_golden("idbn")
    layers = idbn_layers(g, "in_")
    x = T(g["x"]).reshape(g["x"].shape[0], -1)
    seed, bs = int(g["seed"]), int(g["batch"])
    counters = [0, 0]
    for ep in range(int(g["epochs"])):
        for b0 in range(0, x.shape[0], bs):
            flds = [RandomField(seed + li, counters[li]) for li in range(2)]
            O.idbn_train_batch(layers, x[b0:b0 + bs], ep, 1, flds)
            counters = [c + 1 for c in counters]
    for i, st in enumerate(layers):
        check_params(st, g, f"out_l{i}_")
    torch.testing.assert_close(O.idbn_represent(layers, x), T(g["represent"]), **TOL)
    torch.testing.assert_close(O.idbn_represent(layers, x, 1), T(g["represent1"]), **TOL)
    torch.testing.assert_close(O.idbn_reconstruct(layers, x), T(g["reconstruct"]), **TOL)
    torch.testing.assert_close(O.idbn_decode(layers, T(g["decode_in"])), T(g["decode"]), **TOL)


def run_cross(layers, joint, z, y, zc, stream, seed, steps, fe):
    out = O.cross_reconstruct(layers, joint, z, y, steps, z_class_mean=zc, use_free_energy=fe,
                              fld_i2t=RandomField(seed, stream),
                              fld_t2i=RandomField(seed, stream + 1),
                              fld_refine=[RandomField(seed, stream + 2 + c) for c in range(4)])
    return out, stream + 6


def test_imdbn_bias_init_cross_reconstruct_and_train_joint():
    g = load_golden("imdbn")
    K, bs = int(g["K"]), int(g["batch"])
    layers = idbn_layers(g, "in_")
    joint = state_from(g, "in_joint_", [(10, 10 + K)])
    joint.lr, joint.weight_decay, joint.momentum, joint.final_momentum, joint.dynamic_lr = \
        0.04, 1e-4, 0.5, 0.95, True
    x = T(g["x"]).reshape(g["x"].shape[0], -1)
    y = T(g["y"])
    batches = [(x[i:i + bs], y[i:i + bs]) for i in range(0, x.shape[0], bs)]
    zs = [O.idbn_represent(layers, xb) for xb, _ in batches[:2]]
    zc, cnt = O.joint_bias_init(joint, zs, [yb for _, yb in batches[:2]], 10, K)
    torch.testing.assert_close(joint.vb, T(g["bias_vb"]), **TOL)
    torch.testing.assert_close(zc, T(g["z_class_mean"]), **TOL)
    torch.testing.assert_close(cnt, T(g["z_class_count"]), **TOL)
    z = O.idbn_represent(layers, x)
    torch.testing.assert_close(O.hidden_probs(joint, torch.cat([z, y], 1)), T(g["represent"]), **TOL)

    seed = int(g["seed"]) + 100
    (img, py, _, best), stream = run_cross(layers, joint, z, y, zc, 0, seed, 6, False)
    torch.testing.assert_close(img, T(g["cross_img"]), **TOL)
    torch.testing.assert_close(py, T(g["cross_py"]), **TOL)
    assert int(best.max()) == 0          # no free-energy hook -> always the main chain
    (img, py, _, best), stream = run_cross(layers, joint, z, y, zc, stream, seed, 6, True)
    torch.testing.assert_close(img, T(g["cross_img_fe"]), **TOL)
    torch.testing.assert_close(py, T(g["cross_py_fe"]), **TOL)

    # train_joint (imdbn.py:508-639): re-runs the bias init over <=10 batches first
    zs = [O.idbn_represent(layers, xb) for xb, _ in batches]
    zc, _ = O.joint_bias_init(joint, zs, [yb for _, yb in batches], 10, K)
    for ep in range(int(g["train_epochs"])):
        for b, (xb, yb) in enumerate(batches):
            zb = O.idbn_represent(layers, xb)
            B = zb.shape[0]
            vk = torch.zeros(B, 10 + K); km = torch.zeros(B, 10 + K)
            vk[:, 10:] = yb; km[:, 10:] = 1.0
            if ep < 8:
                for _ in range(2):
                    O.cd_train_clamped(joint, vk, km, ep, 1, 4, False, False, True, 0.3, True,
                                       fld=RandomField(seed, stream)); stream += 1
            else:
                O.cd_train(joint, torch.cat([zb, yb], 1), ep, 1, RandomField(seed, stream)); stream += 1
                O.cd_train_clamped(joint, vk, km, ep, 1, 4, False, False, False, 0.3, True,
                                   fld=RandomField(seed, stream)); stream += 1
                if b % 50 == 0:
                    vk2 = torch.zeros(B, 10 + K); km2 = torch.zeros(B, 10 + K)
                    vk2[:, :10] = zb; km2[:, :10] = 1.0
                    O.cd_train_clamped(joint, vk2, km2, ep, 1, 4, False, False, False, 0.3, True,
                                       fld=RandomField(seed, stream)); stream += 1
            _, stream = run_cross(layers, joint, zb, yb, zc, stream, seed, 6, False)
    assert stream == int(g["final_stream"])
    check_params(joint, g, "out_joint_")


def bimodal_states(g):
    mod1 = idbn_layers(g, "in_m1_")
    mod2 = idbn_layers(g, "in_m2_")
    joint = []
    for i in range(2):
        st = state_from(g, f"in_j{i}_")
        st.lr, st.weight_decay, st.momentum, st.final_momentum, st.dynamic_lr = 0.04, 1e-4, 0.5, 0.95, True
        joint.append(st)
    return mod1, mod2, joint


def test_bimodal_bias_init_cross_reconstruct_and_train_joint():
    """iMDBN_BiModal (imdbn_bimodal.py:617-834) restated in the oracle against the reference's own run."""
    g = load_golden("bimodal")
    mod1, mod2, joint = bimodal_states(g)
    x1 = T(g["x1"]).reshape(g["x1"].shape[0], -1)
    x2 = T(g["x2"]).reshape(g["x2"].shape[0], -1)
    bs, seed = int(g["batch"]), int(g["seed"]) + 200
    batches = [(x1[i:i + bs], x2[i:i + bs]) for i in range(0, x1.shape[0], bs)]
    O.bimodal_bias_init(joint[0], [O.idbn_represent(mod1, a) for a, _ in batches[:2]],
                        [O.idbn_represent(mod2, b) for _, b in batches[:2]], 12)
    torch.testing.assert_close(joint[0].vb, T(g["bias_vb"]), **TOL)
    z1, z2 = O.idbn_represent(mod1, x1), O.idbn_represent(mod2, x2)
    h = torch.cat([z1, z2], 1)
    for st in joint:
        h = O.hidden_probs(st, h)
    torch.testing.assert_close(h, T(g["represent"]), **TOL)
    r1, r2 = O.bimodal_cross_reconstruct(mod1, mod2, joint[0], z1, z2, 5, seed, 0)
    torch.testing.assert_close(r1, T(g["cross_mod1"]), **TOL)
    torch.testing.assert_close(r2, T(g["cross_mod2"]), **TOL)

    # train_joint continues the call numbering of joint layer 0 after the two chains above
    hist = O.bimodal_train_joint(mod1, mod2, joint, batches, int(g["train_epochs"]), joint_cd=2, aux_steps=4,
                                 cross_steps=5, seeds=[seed, seed + 1], bias_batches=10, streams=[2, 0])
    for i, st in enumerate(joint):
        check_params(st, g, f"out_j{i}_")
    np.testing.assert_allclose([h["mod1_mse"] for h in hist], g["mod1_mse"], rtol=1e-5)
    np.testing.assert_allclose([h["mod2_mse"] for h in hist], g["mod2_mse"], rtol=1e-5)
    np.testing.assert_allclose([h["cd_loss"] for h in hist if h["cd_loss"] == h["cd_loss"]], g["cd_loss"], rtol=1e-5)


TRACE_LISTS = ("p_top1", "p_top2", "p_gap", "p_gt", "deltaF_pred_traj")
TRACE_SCALARS = ("steps_to_converge", "kstar", "predT", "margin_energy", "fe_top1_final", "fe_gap_final", "gt")


def test_energy_diagnostics():
    """utils/energy_utils.py: class_free_energies and trace_single_img2txt against the reference's own run."""
    g = load_golden("energy")
    K = int(g["K"])
    layers = [state_from(g, f"l{i}_") for i in range(2)]
    joint = state_from(g, "joint_", [(12, 12 + K)])
    x = T(g["x"]).reshape(g["x"].shape[0], -1)
    z = O.idbn_represent(layers, x)
    torch.testing.assert_close(O.class_free_energies(joint, z, K, 12), T(g["Fk"]), rtol=1e-5, atol=1e-5)
    # F_k is the free energy of [z, e_k]
    ek = torch.eye(K)
    brute = torch.stack([O.free_energy(joint, torch.cat([z, ek[k].expand(z.shape[0], K)], 1)) for k in range(K)], 1)
    torch.testing.assert_close(O.class_free_energies(joint, z, K, 12), brute, rtol=1e-5, atol=1e-5)
    for i in range(4):
        tr = O.trace_single_img2txt(layers, joint, x[i:i + 1], int(g[f"t{i}_gt"]), K, steps=12, eps_l1=1e-4,
                                    stable_steps=3, gap_thresh=0.9)
        for key in TRACE_LISTS:
            np.testing.assert_allclose(tr[key], g[f"t{i}_{key}"], rtol=1e-5, atol=1e-6, err_msg=key)
        for key in TRACE_SCALARS:
            np.testing.assert_allclose(tr[key], g[f"t{i}_{key}"], rtol=1e-5, atol=1e-6, err_msg=key)


def test_rbm_extra_backward_sample_gibbs_step_annealed():
    """backward_sample, gibbs_step and conditional_gibbs_annealed (rbm.py:153-178, 240-298) against the reference."""
    g = load_golden("rbm_extra")
    st = state_from(g, "in_", g["groups"])
    seed = int(g["seed"])
    h, v0 = T(g["h"]), T(g["v0"])
    assert torch.equal(O.backward_sample(st, h, RandomField(seed, 0)), T(g["backward_sample"]))
    for i, (sh, sv) in enumerate(g["gs_cfg"]):
        vn, vp, hh, hp = O.gibbs_step(st, v0, bool(sh), bool(sv), RandomField(seed, 1 + i))
        torch.testing.assert_close(vn, T(g[f"gs{i}_v_next"]), **TOL)
        torch.testing.assert_close(vp, T(g[f"gs{i}_v_prob"]), **TOL)
        torch.testing.assert_close(hh, T(g[f"gs{i}_h"]), **TOL)
        torch.testing.assert_close(hp, T(g[f"gs{i}_h_prob"]), **TOL)
    vk, km = T(g["v_known"]), T(g["km"])
    for i, (n, T0, T1, until, every, final) in enumerate(g["cga_cfg"]):
        out = O.conditional_gibbs_annealed(st, vk, km, int(n), float(T0), float(T1), int(until), int(every),
                                           bool(final), RandomField(seed, 10 + i))
        torch.testing.assert_close(out, T(g[f"cga{i}"]), **TOL)


def test_finetune_image_last_layer():
    """iMDBN.finetune_image_last_layer (imdbn.py:344-384): only the last image layer moves, lr restored."""
    g = load_golden("finetune")
    layers = idbn_layers(g, "in_")
    x = T(g["x"]).reshape(g["x"].shape[0], -1)
    seed, bs = int(g["seed"]), int(g["batch"])
    batches = [x[b0:b0 + bs] for b0 in range(0, x.shape[0], bs)]
    lr0 = layers[-1].lr
    O.finetune_last_layer(layers, batches, 2, 0.3, 2, [RandomField(seed + 1, s) for s in range(2 * len(batches))])
    assert layers[-1].lr == lr0
    for i, st in enumerate(layers):
        check_params(st, g, f"out_l{i}_")
    torch.testing.assert_close(layers[0].W, T(g["in_l0_W"]))
