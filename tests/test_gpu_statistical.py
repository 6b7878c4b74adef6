"""North-star parity bullet 3: UNSEEDED runs must give statistically matched curves.  A small iMDBN is
trained on a learnable synthetic task (label = number of bright quadrants) by the CPU oracle and by the
CUDA path with DIFFERENT random fields, in both precisions; reconstruction loss and IMG->TXT label
accuracy must land in the same band."""
import numpy as np
import pytest
import torch

from oracle import rbm_oracle as O
from oracle.philox import RandomField

pytestmark = pytest.mark.gpu
DEV = "cuda"
D, H1, HJ, K, N, BS, EPOCHS = 64, 32, 24, 4, 256, 32, 12


def dataset(seed):
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, K, (N,), generator=g)
    x = torch.zeros(N, D)
    for i in range(N):
        p = torch.full((D,), 0.08)
        p[int(y[i]) * (D // K):(int(y[i]) + 1) * (D // K)] = 0.85      # one bright block per class
        x[i] = (torch.rand(D, generator=g) < p).float()
    return x, torch.nn.functional.one_hot(y, K).float()


def run_oracle(x, y, seed):
    l0 = O.new_state(D, H1, seed=seed, lr=0.1, weight_decay=1e-4, momentum=0.5, final_momentum=0.95, dynamic_lr=True)
    jr = O.new_state(H1 + K, HJ, seed=seed + 1, groups=[(H1, H1 + K)], lr=0.1, weight_decay=1e-4, momentum=0.5,
                     final_momentum=0.95, dynamic_lr=True)
    s, losses = 0, []
    for ep in range(EPOCHS):
        for b in range(0, N, BS):
            loss, _ = O.cd_train(l0, x[b:b + BS], ep, 1, RandomField(seed, s)); s += 1
            losses.append(float(loss))
    for ep in range(EPOCHS):
        for b in range(0, N, BS):
            z = O.hidden_probs(l0, x[b:b + BS])
            O.cd_train(jr, torch.cat([z, y[b:b + BS]], 1), ep, 1, RandomField(seed, s)); s += 1
    z = O.hidden_probs(l0, x)
    vk = torch.zeros(N, H1 + K); km = torch.zeros(N, H1 + K); vk[:, :H1] = z; km[:, :H1] = 1
    py = O.conditional_gibbs(jr, vk, km, n_steps=20, fld=RandomField(seed, s))[:, H1:]
    return np.mean(losses[-8:]), float((py.argmax(1) == y.argmax(1)).float().mean())


def run_gpu(M, x, y, seed):
    torch.manual_seed(seed)
    l0 = M.RBM(D, H1, 0.1, 1e-4, 0.5, dynamic_lr=True, final_momentum=0.95).to(DEV)
    jr = M.RBM(H1 + K, HJ, 0.1, 1e-4, 0.5, dynamic_lr=True, final_momentum=0.95,
               softmax_groups=[(H1, H1 + K)]).to(DEV)
    xd, yd = x.to(DEV), y.to(DEV)
    losses = []
    for ep in range(EPOCHS):
        for b in range(0, N, BS):
            losses.append(l0.train_epoch(xd[b:b + BS], ep, EPOCHS, CD=1))
    for ep in range(EPOCHS):
        for b in range(0, N, BS):
            z = l0.forward(xd[b:b + BS])
            jr.train_epoch(torch.cat([z, yd[b:b + BS]], 1), ep, EPOCHS, CD=1)
    z = l0.forward(xd)
    vk = torch.zeros(N, H1 + K, device=DEV); km = torch.zeros_like(vk); vk[:, :H1] = z; km[:, :H1] = 1
    py = jr.conditional_gibbs(vk, km, n_steps=20, clamp_prefix=H1)[:, H1:]
    return float(torch.stack(losses[-8:]).mean()), float((py.argmax(1) == yd.argmax(1)).float().mean())


def test_unseeded_curves_match_statistically():
    import multimodal_idbn_b200 as M
    x, y = dataset(0)
    ref = [run_oracle(x, y, s) for s in (1, 2, 3)]
    ref_loss, ref_acc = np.mean([r[0] for r in ref]), np.mean([r[1] for r in ref])
    assert ref_acc > 0.8, ref                      # the task is learnable by the reference algorithm
    for prec in ("fp32", "tf32"):
        M.set_precision(prec)
        try:
            got = [run_gpu(M, x, y, s) for s in (11, 12, 13)]
        finally:
            M.set_precision("fp32")
        loss, acc = np.mean([g[0] for g in got]), np.mean([g[1] for g in got])
        assert abs(loss - ref_loss) < 0.15 * ref_loss + 2e-3, (prec, got, ref)
        assert acc > ref_acc - 0.08, (prec, got, ref)
