"""GPU tests of the EXACT tensor-core mode (precision 'tf32x2'): tcgen05.mma kind::tf32 with every operand entering
as hi + lo terms (csrc/tc_gemm.cu, tc_stats.cuh), IEEE finishes.  It has to meet the SAME bars as the fp32 FFMA engine
(north_star: sampled states bit-exact outside a 1e-6 band, activations / weights to fp32 rounding): the whole of
tests/test_gpu_parity.py runs in this mode too; this file adds the shapes, batch sizes and operand kinds that only the
tensor-core kernels see (stream-K edges, two-source batches, chunked batches, real-valued operands on both sides)."""
import ctypes as C

import pytest
import torch

from oracle import rbm_oracle as O
from oracle.philox import RandomField

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = dict(rtol=2e-5, atol=2e-6)


@pytest.fixture()
def M():
    import multimodal_idbn_b200 as m
    m.load_library()
    m.set_precision("tf32x2")
    yield m
    m.set_precision("fp32")


def make(M, V, H, seed=0, scale=1.0, groups=None):
    st = O.new_state(V, H, seed=seed, groups=groups, lr=0.1, weight_decay=1e-4, momentum=0.5,
                     final_momentum=0.95, dynamic_lr=True)
    st.W *= scale
    g = torch.Generator().manual_seed(seed + 1)
    st.hb.copy_(torch.randn(H, generator=g) * 0.1); st.vb.copy_(torch.randn(V, generator=g) * 0.1)
    r = M.RBM(V, H, 0.1, 1e-4, 0.5, dynamic_lr=True, final_momentum=0.95,
              softmax_groups=list(groups or [])).to(DEV)
    with torch.no_grad():
        r.W.data.copy_(st.W); r.hid_bias.data.copy_(st.hb); r.vis_bias.data.copy_(st.vb)
    return st, r


SHAPES = [(10000, 1500), (1500, 500), (532, 256), (128, 128), (36, 20), (2080, 1024)]


@pytest.mark.parametrize("V,H", SHAPES)
@pytest.mark.parametrize("B", [1, 64, 130, 256, 700])
def test_up_down_exact_vs_oracle(M, V, H, B):
    st, r = make(M, V, H, seed=V + H)
    gen = torch.Generator().manual_seed(B)
    v = (torch.rand(B, V, generator=gen) < 0.3).float()
    h = torch.rand(B, H, generator=gen)
    torch.testing.assert_close(r.forward(v.to(DEV)).cpu(), O.hidden_probs(st, v), **TOL)
    torch.testing.assert_close(r.visible_probs(h.to(DEV)).cpu(), O.visible_probs(st, h), **TOL)
    # real-valued (non-binary) inputs on both sides, as test_extraction.py feeds (randn): |x| up to ~4, so the
    # pre-activations are sums of 10^4 terms of magnitude 1e-2: absolute fp32 rounding of the sum, not of the output
    x = torch.randn(B, V, generator=gen)
    torch.testing.assert_close(r.forward(x.to(DEV)).cpu(), O.hidden_probs(st, x), rtol=2e-5, atol=5e-6)
    torch.testing.assert_close(r.forward(x.to(DEV), T=0.7).cpu(), O.hidden_probs(st, x, T=0.7), rtol=2e-5, atol=5e-6)


@pytest.mark.parametrize("V,H", SHAPES)
@pytest.mark.parametrize("B", [1, 64, 100, 600])
@pytest.mark.parametrize("real_v", [False, True])
def test_assoc_stats_exact(M, V, H, B, real_v):
    from multimodal_idbn_b200 import _lib as L
    st, r = make(M, V, H, seed=3)
    gen = torch.Generator().manual_seed(7)
    if real_v:
        vp = torch.rand(B, V, generator=gen); vn = torch.rand(B, V, generator=gen)
    else:
        vp = (torch.rand(B, V, generator=gen) < 0.3).float(); vn = (torch.rand(B, V, generator=gen) < 0.3).float()
    hp = torch.rand(B, H, generator=gen); hn = torch.rand(B, H, generator=gen)
    ref = vp.double().T @ hp.double() - vn.double().T @ hn.double()
    out = torch.empty(V, H, device=DEV)
    ctx, stream = L.context_for(out)
    rs = r._struct()
    d = [t.to(DEV).contiguous() for t in (vp, hp, vn, hn)]
    ctx.check(ctx.lib.imdbn_assoc_stats(ctx.handle, C.byref(rs), L.ptr(d[0]), L.ptr(d[1]), L.ptr(d[2]),
                                        L.ptr(d[3]), B, L.ptr(out), stream), "assoc")
    # fp32 accumulation of B products of magnitude <= 1 (the two phases cancel): absolute error ~ B * 2^-24 * few
    # (the tensor core's accumulator truncates instead of rounding: a few times the bound of a rounded fp32 sum)
    torch.testing.assert_close(out.cpu().double(), ref, rtol=2e-6, atol=(2e-7 if B <= 100 else 1e-6) * max(8, B))


def test_cd10_wide_layer_vs_oracle(M):
    """CD-10 on a 2080 -> 1024 RBM with a 32-way softmax group (the C5 joint shape), batch 128: loss and update
    against the oracle at the fp32 bars."""
    V, H, Dz, B = 2080, 1024, 2048, 128
    st, r = make(M, V, H, seed=13, scale=2.0, groups=[(Dz, V)])
    data = torch.cat([O.synthetic_images(B, Dz, p=0.3, seed=5), O.synthetic_labels(B, V - Dz, seed=6)], 1)
    loss_ref, _ = O.cd_train(st, data, 0, 10, RandomField(9, 0))
    r.set_rng(9, 0)
    loss = r.train_epoch(data.to(DEV), 0, 1, CD=10)
    torch.testing.assert_close(loss.cpu(), loss_ref, rtol=1e-4, atol=1e-6)
    # a unit whose probability is within rounding of its uniform may flip and move one row / column of dS by lr/B
    bad = ((r.W.detach().cpu() - st.W).abs() > 2e-6 + 2e-5 * st.W.abs()).float().mean()
    assert float(bad) < 2e-3, float(bad)


@pytest.mark.parametrize("prec", ["tf32", "tf32x2"])
def test_clamped_cd_large_batch_vs_oracle(prec):
    """train_epoch_clamped at a batch that takes the STEPPED tensor-core chain (B >= 512): the positive phase must be
    the noisy mean-field inference, not a copy of v_known (regression test for a zero-initialised block-mask hint)."""
    import multimodal_idbn_b200 as M
    M.set_precision(prec)
    try:
        V, H, Dz, K, B = 532, 256, 500, 32, 640
        st, r = make(M, V, H, seed=17, scale=3.0, groups=[(Dz, V)])
        y = O.synthetic_labels(B, K, seed=3)
        vk = torch.zeros(B, V); km = torch.zeros(B, V); vk[:, Dz:] = y; km[:, Dz:] = 1
        W0 = st.W.clone()
        loss_ref, _ = O.cd_train_clamped(st, vk, km, 0, k=1, cond_init_steps=12, sample_h=False, sample_v=False,
                                         reclamp_negative=False, aux_lr_mult=0.3, use_noisy_init=True,
                                         fld=RandomField(23, 0))
        r.set_rng(23, 0)
        loss = r.train_epoch_clamped(vk.to(DEV), km.to(DEV), 0, 1, CD=1, cond_init_steps=12, sample_h=False,
                                     sample_v=False, reclamp_negative=False, aux_lr_mult=0.3, use_noisy_init=True)
        tol = dict(rtol=2e-2, atol=1e-4) if prec == "tf32" else dict(rtol=2e-4, atol=1e-6)
        torch.testing.assert_close(loss.cpu(), torch.as_tensor(loss_ref, dtype=torch.float32), **tol)
        dW_ref = st.W - W0
        dW = r.W.detach().cpu() - W0
        assert float(dW_ref.abs().max()) > 0
        rel = float((dW - dW_ref).abs().max() / dW_ref.abs().max())
        assert rel < (5e-2 if prec == "tf32" else 2e-4), rel
    finally:
        M.set_precision("fp32")


def test_large_batch_chains_exact(M):
    """B >= 512 mean-field chains run step by step on the tensor-core passes; in exact mode they must agree with the
    oracle like the persistent fp32 chain kernel does."""
    V, H, Dz, K, B = 532, 256, 500, 32, 600
    st, r = make(M, V, H, seed=21, scale=2.0, groups=[(Dz, V)])
    z = torch.rand(B, Dz, generator=torch.Generator().manual_seed(2))
    y = O.synthetic_labels(B, K, seed=3)
    mu = torch.rand(B, Dz, generator=torch.Generator().manual_seed(4))
    vk = torch.zeros(B, V); km = torch.zeros(B, V); vk[:, Dz:] = y; km[:, Dz:] = 1
    ref = O.noisy_meanfield(st, vk, km, n_steps=20, mu_pull=(mu, 0.15), fld=RandomField(31, 0))
    r._mu_pull = {"mu_k": mu.to(DEV), "eta0": 0.15}
    r.set_rng(31, 0)
    out = r.noisy_meanfield_annealed(vk.to(DEV), km.to(DEV), n_steps=20)
    r._mu_pull = None
    torch.testing.assert_close(out.cpu(), ref, rtol=2e-4, atol=3e-5)
    vk = torch.zeros(B, V); km = torch.zeros(B, V); vk[:, :Dz] = z; km[:, :Dz] = 1
    ref = O.conditional_gibbs(st, vk, km, n_steps=20, fld=RandomField(31, 1))
    r.set_rng(31, 1)
    out = r.conditional_gibbs(vk.to(DEV), km.to(DEV), n_steps=20)
    torch.testing.assert_close(out.cpu(), ref, rtol=2e-4, atol=3e-5)


def test_exact_mode_runs_on_the_tensor_core_kernels(M):
    """The exact mode must not fall back to the FFMA engine: the in-library profile sees tensor-core-sized timings and
    the launch count of a CD-1 update equals the tf32 mode's."""
    import multimodal_idbn_b200 as mm
    V, H, B = 1500, 500, 64
    counts = {}
    for mode in ("tf32", "tf32x2"):
        mm.set_precision(mode)
        _, r = make(mm, V, H, seed=1)
        data = O.synthetic_images(B, V, seed=2).to(DEV)
        r.train_epoch(data, 0, 1, CD=1)
        torch.cuda.synchronize()
        n0 = mm.total_launches()
        r.train_epoch(data, 0, 1, CD=1)
        torch.cuda.synchronize()
        counts[mode] = mm.total_launches() - n0
    assert counts["tf32"] == counts["tf32x2"], counts
