"""The single-kernel CD update of small layers (csrc/cd_small.cuh: upper iDBN layers, V*H <= 2M, batch <= 64,
no softmax groups) against the oracle: plain / sparse / CD-3 / ragged batch, and the fused
train + forward + next-batch positive phase that iDBN.train_step drives (reference rbm.py:180-227, idbn.py:199-204)."""
import pytest
import torch

from oracle import rbm_oracle as O
from oracle.philox import RandomField

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = dict(rtol=1e-4, atol=2e-5)


@pytest.fixture(scope="module")
def M():
    import multimodal_idbn_b200 as M
    M.set_precision("fp32")          # the single-kernel path is the fp32-mode implementation of small layers
    return M


def pair(M, V, H, seed, **hyper):
    st = O.new_state(V, H, seed=seed, **hyper)
    st.W *= 2.0
    g = torch.Generator().manual_seed(seed + 1)
    st.hb.copy_(torch.randn(H, generator=g) * 0.1); st.vb.copy_(torch.randn(V, generator=g) * 0.1)
    r = M.RBM(V, H, st.lr, st.weight_decay, st.momentum, dynamic_lr=st.dynamic_lr, final_momentum=st.final_momentum,
              sparsity=st.sparsity, sparsity_factor=st.sparsity_factor).to(DEV)
    with torch.no_grad():
        r.W.data.copy_(st.W); r.hid_bias.data.copy_(st.hb); r.vis_bias.data.copy_(st.vb)
    return st, r


def check_state(r, st, flips_ok=0.002):
    dW = (r.W.detach().cpu() - st.W).abs()
    # a unit inside the 1e-6 band of its threshold may flip; everything else must be tight
    assert float((dW > 2e-5).float().mean()) <= flips_ok, float(dW.max())
    torch.testing.assert_close(r.hid_bias.detach().cpu(), st.hb, rtol=1e-3, atol=2e-3 if flips_ok else 2e-5)
    torch.testing.assert_close(r.vis_bias.detach().cpu(), st.vb, rtol=1e-3, atol=2e-3 if flips_ok else 2e-5)


@pytest.mark.parametrize("V,H,B,k,sparse", [(1500, 500, 64, 1, False), (1500, 500, 64, 3, True), (1500, 500, 37, 1, False),
                                            (100, 36, 5, 2, True), (2048, 1024, 64, 1, False), (64, 64, 1, 1, False)])
def test_cd_small_matches_oracle(M, V, H, B, k, sparse):
    hyper = dict(lr=0.1, weight_decay=1e-4, momentum=0.5, final_momentum=0.95, dynamic_lr=True, sparsity=sparse,
                 sparsity_factor=0.1)
    st, r = pair(M, V, H, 21, **hyper)
    data = torch.rand(B, V, generator=torch.Generator().manual_seed(9))          # layer >= 1 inputs are probabilities
    for step in range(3):
        loss_ref, _ = O.cd_train(st, data, 7, k, RandomField(5, step))
        r.set_rng(5, step)
        loss = r.train_epoch(data.to(DEV), 7, 10, CD=k)
        torch.testing.assert_close(loss.cpu(), torch.as_tensor(loss_ref, dtype=torch.float32), rtol=1e-4, atol=1e-6)
        check_state(r, st)
        torch.testing.assert_close(r.W_m.cpu(), st.Wm, rtol=1e-3, atol=2e-3)
        # resynchronise so that a legitimate band flip of one step does not leak into the next comparison
        with torch.no_grad():
            r.W.data.copy_(st.W); r.hid_bias.data.copy_(st.hb); r.vis_bias.data.copy_(st.vb)
            r.W_m.copy_(st.Wm); r.hb_m.copy_(st.hbm); r.vb_m.copy_(st.vbm)


def test_cd_small_fused_forward_and_next_positive_phase(M):
    V, H, B = 1500, 500, 64
    hyper = dict(lr=0.1, weight_decay=1e-4, momentum=0.5, final_momentum=0.95, dynamic_lr=True)
    st, r = pair(M, V, H, 4, **hyper)
    g = torch.Generator().manual_seed(3)
    batches = [torch.rand(B, V, generator=g) for _ in range(4)]
    dev = [b.to(DEV) for b in batches]
    r.set_rng(8, 0)
    for i in range(3):
        loss_ref, _ = O.cd_train(st, batches[i], 0, 1, RandomField(8, i))
        h_ref = O.hidden_probs(st, batches[i])
        loss, h = r.train_epoch_fwd(dev[i], 0, 1, CD=1, next_data=dev[i + 1])
        torch.testing.assert_close(loss.cpu(), torch.as_tensor(loss_ref, dtype=torch.float32), rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(h.cpu(), h_ref, rtol=1e-3, atol=5e-4)
        pos_next = r._pos_cache[1]
        torch.testing.assert_close(pos_next.cpu(), O.hidden_probs(st, batches[i + 1]), rtol=1e-3, atol=5e-4)
        check_state(r, st)
        with torch.no_grad():
            r.W.data.copy_(st.W); r.hid_bias.data.copy_(st.hb); r.vis_bias.data.copy_(st.vb)
            r.W_m.copy_(st.Wm); r.hb_m.copy_(st.hbm); r.vb_m.copy_(st.vbm)
            r._pos_cache = None                                  # parameters were overwritten: drop the cache
    # with the cache live (no resync) the cached positive phase must equal a fresh up pass
    r.set_rng(8, 10)
    _, _ = r.train_epoch_fwd(dev[0], 0, 1, CD=1, next_data=dev[1])
    W1 = r.W.detach().clone()
    cached = r._pos_cache[1].clone()
    torch.testing.assert_close(cached, r.forward(dev[1]), rtol=1e-4, atol=1e-5)
    loss2, _ = r.train_epoch_fwd(dev[1], 0, 1, CD=1, next_data=dev[2])      # consumes the cache
    assert torch.isfinite(loss2) and not torch.equal(W1, r.W.detach())
